// Launch entry points of the nearest-neighbour kernels (nn_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstddef>

namespace mplb {

// number of key slices the scan is split into (>= 1); partial buffers are [nq][slices][k]
int nnSlices(size_t nq, size_t n, int numSms);

cudaError_t launchNnToFloat(const double *dSrc, float *dDst, size_t n, cudaStream_t stream);
// dOut[2] = {largest |translation coordinate| (SE3) / |coordinate| (L1) rounded up -- NaN/inf if a key is not finite --,
// 1.0f if an SE3 key's quaternion is not unit} of n keys in device memory
cudaError_t launchNnKeyStats(const double *dKeys, size_t n, int dim, int metric, float *dOut, cudaStream_t stream);

// metric 0: SE3 (7-double keys), 1: L1 (dim doubles).  dDist/dIdx: [nq][k].  dKeys32 is the FP32 copy of the
// keys, keysAbsMax the largest |translation coordinate| (SE3) / |coordinate| (L1) among them; dPartVal
// [nq][slices][k] floats, dThr [nq] floats, dPartDist/dPartIdx [nq][slices][k] (unused when slices == 1).
cudaError_t launchNnSearch(const double *dKeys, const float *dKeys32, float keysAbsMax, size_t n, int dim, int metric,
                           double so3w, double l2w, const double *dQueries, size_t nq, int k, int slices,
                           float *dPartVal, float *dThr, double *dPartDist, long long *dPartIdx, double *dDist,
                           long long *dIdx, unsigned long long *dExactCount, cudaStream_t stream);

// ---- spatial index over the SE(3) keys' translations (nn_grid.cu)
struct NnGridGeom {
    double origin[3];    // lower corner of the grid
    double cell, invCell;
    int dims[3];         // cells per axis (x fastest in memory)
    int metric;          // 0: SE3 keys (7 doubles; the grid is over the translation, coordinates 4..6), 1: L1 keys
    int dim, coord0;     // doubles per key; first of the three key coordinates the grid is laid over (L1: 0..2)
};
// {min x, y, z, max x, y, z} of the translations of n keys as order-preserving bit patterns (nnGridOrderedToFloat)
cudaError_t launchNnGridBounds(const double *dKeys, size_t n, int dim, int coord0, unsigned *dOut6, cudaStream_t stream);
float nnGridOrderedToFloat(unsigned bits);
// counting sort of keys [0, n) by cell: dStart [cells + 1], dPerm [n] (sorted position -> key index), dSorted [n][2]
// float4 (FP32 copy in sorted order); dCellOf [n] and dCounts [cells] are scratch
cudaError_t launchNnGridBuild(const double *dKeys, size_t n, const NnGridGeom &g, unsigned *dCellOf, unsigned *dCounts,
                              unsigned *dStart, unsigned *dPerm, float4 *dSorted, cudaStream_t stream);
// exact k nearest among the nBuilt indexed keys, one warp per query: dDist / dIdx [nq][k] ascending by (distance,
// index), padded with (+inf, -1); dStats (optional): {FP64 evaluations, queries that took the slow selection}
cudaError_t launchNnGridSearch(const NnGridGeom &g, const unsigned *dStart, const unsigned *dPerm, const float4 *dSorted,
                               const double *dKeys, size_t nBuilt, double so3w, double l2w, float keysAbsMax,
                               const double *dQueries, size_t nq, int k, double *dDist, long long *dIdx,
                               unsigned long long *dStats, cudaStream_t stream);
// merges two ascending lists per query (B's indices are relative to offB)
cudaError_t launchNnMerge2(const double *dA, const long long *iA, const double *dB, const long long *iB, long long offB,
                           size_t nq, int k, double *outD, long long *outI, cudaStream_t stream);

}  // namespace mplb
