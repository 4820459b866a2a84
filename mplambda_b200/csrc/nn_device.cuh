// Launch entry points of the nearest-neighbour kernels (nn_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstddef>

namespace mplb {

// number of key slices the scan is split into (>= 1); partial buffers are [nq][slices][k]
int nnSlices(size_t nq, size_t n, int numSms);

// metric 0: SE3 (7-double keys), 1: L1 (dim doubles).  dDist/dIdx: [nq][k]
cudaError_t launchNnSearch(const double *dKeys, size_t n, int dim, int metric, double so3w, double l2w,
                           const double *dQueries, size_t nq, int k, int slices, double *dPartDist,
                           long long *dPartIdx, double *dDist, long long *dIdx, cudaStream_t stream);

}  // namespace mplb
