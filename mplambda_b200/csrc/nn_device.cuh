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

}  // namespace mplb
