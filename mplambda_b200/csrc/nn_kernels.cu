// Batched exact nearest-neighbour kernels for sm_100a.
//
// Replaces the query side of nigh::Nigh as the planners use it (reference include/mpl/prrt.hpp:66-68
// `nn_.nearest(scale(q))`, include/mpl/pcforest.hpp:80-87 `nn_.nearest(nbh, scale(q), k)`): exact
// 1-NN / k-NN under nigh's SE3Space<S,a,b> metric  a*acos(|q.q'|) + b*||t-t'||  or L1Space<S,dim>
// ([EXT] nigh), ties resolved to the lowest insertion index, k-NN sorted ascending.
//
// Brute force, tiled: a CTA owns 128 queries (one per thread, key in registers) and one slice of
// the key set; key tiles are staged through shared memory with coalesced 128-bit loads and every
// thread scans the tile (shared-memory broadcast reads).  Distances are evaluated in FP64 with the
// oracle's operation order (no FMA contraction), so the L1 metric is bit-identical to the CPU
// path and the SE3 metric differs at most in acos' last bit.  Slices give the grid enough CTAs
// when there are few queries; a second kernel merges the per-slice candidate lists.
#include "nn_device.cuh"

#include <cfloat>

namespace mplb {
namespace {

constexpr int kNnThreads = 128;
constexpr int kNnTile = 128;   // keys per shared-memory tile
constexpr int kNnMaxK = 64;
constexpr int kNnMaxDim = 8;

__device__ __forceinline__ double nnDistance(const double *q, const double *key, int dim, int metric, double so3w,
                                             double l2w) {
    if (metric == 0) {
        double d = fabs(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(key[0], q[0]), __dmul_rn(key[1], q[1])),
                                            __dmul_rn(key[2], q[2])),
                                  __dmul_rn(key[3], q[3])));
        double so3 = d < 1.0 ? acos(d) : 0.0;
        double dx = __dsub_rn(key[4], q[4]), dy = __dsub_rn(key[5], q[5]), dz = __dsub_rn(key[6], q[6]);
        double l2 = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        return __dadd_rn(__dmul_rn(so3w, so3), __dmul_rn(l2w, l2));
    }
    double s = 0;
    for (int i = 0; i < dim; ++i) s = __dadd_rn(s, fabs(__dsub_rn(key[i], q[i])));
    return s;
}

// grid = (query blocks, slices).  out*: [nq][slices][k]
__global__ void __launch_bounds__(kNnThreads)
nnScanKernel(const double *__restrict__ keys, unsigned long long n, int dim, int metric, double so3w, double l2w,
             const double *__restrict__ queries, unsigned long long nq, int k, unsigned long long keysPerSlice,
             double *__restrict__ outDist, long long *__restrict__ outIdx) {
    __shared__ __align__(16) double tile[kNnTile * kNnMaxDim];
    const unsigned long long qi = blockIdx.x * (unsigned long long)kNnThreads + threadIdx.x;
    const bool haveQuery = qi < nq;
    double q[kNnMaxDim];
#pragma unroll
    for (int i = 0; i < kNnMaxDim; ++i) q[i] = (haveQuery && i < dim) ? __ldg(queries + qi * dim + i) : 0.0;

    double bestD[kNnMaxK];
    long long bestI[kNnMaxK];
    int cnt = 0;
    double d1 = INFINITY;   // k == 1 fast path
    long long i1 = -1;

    const unsigned long long lo = blockIdx.y * keysPerSlice;
    const unsigned long long hi = min(n, lo + keysPerSlice);
    for (unsigned long long t0 = lo; t0 < hi; t0 += kNnTile) {
        const int tn = (int)min((unsigned long long)kNnTile, hi - t0);
        const int nd = tn * dim;
        __syncthreads();
        {   // coalesced tile load (pairs of doubles when aligned)
            const double *src = keys + t0 * dim;
            for (int i = threadIdx.x; i < nd; i += kNnThreads) tile[i] = __ldg(src + i);
        }
        __syncthreads();
        if (!haveQuery) continue;
        if (k == 1) {
            for (int j = 0; j < tn; ++j) {
                const double d = nnDistance(q, tile + j * dim, dim, metric, so3w, l2w);
                if (d < d1) {
                    d1 = d;
                    i1 = (long long)(t0 + j);
                }
            }
        } else {
            for (int j = 0; j < tn; ++j) {
                const double d = nnDistance(q, tile + j * dim, dim, metric, so3w, l2w);
                if (cnt == k && !(d < bestD[k - 1])) continue;
                int pos = cnt < k ? cnt : k - 1;
                while (pos > 0 && bestD[pos - 1] > d) {   // equal distances keep their (index) order
                    bestD[pos] = bestD[pos - 1];
                    bestI[pos] = bestI[pos - 1];
                    --pos;
                }
                bestD[pos] = d;
                bestI[pos] = (long long)(t0 + j);
                if (cnt < k) ++cnt;
            }
        }
    }
    if (!haveQuery) return;
    double *od = outDist + (qi * gridDim.y + blockIdx.y) * k;
    long long *oi = outIdx + (qi * gridDim.y + blockIdx.y) * k;
    if (k == 1) {
        od[0] = d1;
        oi[0] = i1;
    } else {
        for (int r = 0; r < k; ++r) {
            od[r] = r < cnt ? bestD[r] : INFINITY;
            oi[r] = r < cnt ? bestI[r] : -1;
        }
    }
}

// one thread per query: merge `slices` ascending lists of length k into one (ties -> lower index)
__global__ void nnMergeKernel(const double *__restrict__ inDist, const long long *__restrict__ inIdx,
                              unsigned long long nq, int slices, int k, double *__restrict__ outDist,
                              long long *__restrict__ outIdx) {
    const unsigned long long qi = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const double *d = inDist + qi * slices * k;
    const long long *ix = inIdx + qi * slices * k;
    unsigned char head[256];
    for (int s = 0; s < slices; ++s) head[s] = 0;
    for (int r = 0; r < k; ++r) {
        int bs = -1;
        double bd = INFINITY;
        long long bi = -1;
        for (int s = 0; s < slices; ++s) {
            if (head[s] >= k) continue;
            const long long ci = ix[s * k + head[s]];
            if (ci < 0) continue;
            const double cd = d[s * k + head[s]];
            if (bs < 0 || cd < bd || (cd == bd && ci < bi)) {
                bs = s;
                bd = cd;
                bi = ci;
            }
        }
        outDist[qi * k + r] = bs < 0 ? INFINITY : bd;
        outIdx[qi * k + r] = bs < 0 ? -1 : bi;
        if (bs >= 0) ++head[bs];
    }
}

}  // namespace

int nnSlices(size_t nq, size_t n, int numSms) {
    const size_t qBlocks = (nq + kNnThreads - 1) / kNnThreads;
    size_t want = (size_t)(4 * numSms + qBlocks - 1) / qBlocks;           // aim for ~4 CTAs per SM
    const size_t maxByKeys = (n + 4 * kNnTile - 1) / (4 * kNnTile);          // at least 4 tiles per slice
    want = std::max<size_t>(1, std::min(want, std::max<size_t>(1, maxByKeys)));
    return (int)std::min<size_t>(want, 256);
}

cudaError_t launchNnSearch(const double *dKeys, size_t n, int dim, int metric, double so3w, double l2w,
                           const double *dQueries, size_t nq, int k, int slices, double *dPartDist,
                           long long *dPartIdx, double *dDist, long long *dIdx, cudaStream_t stream) {
    if (nq == 0) return cudaSuccess;
    const size_t keysPerSlice = ((n + slices - 1) / slices + kNnTile - 1) / kNnTile * kNnTile;
    dim3 grid((unsigned)((nq + kNnThreads - 1) / kNnThreads), (unsigned)slices);
    if (slices == 1) {
        nnScanKernel<<<grid, kNnThreads, 0, stream>>>(dKeys, n, dim, metric, so3w, l2w, dQueries, nq, k,
                                                      std::max<size_t>(keysPerSlice, kNnTile), dDist, dIdx);
        return cudaGetLastError();
    }
    nnScanKernel<<<grid, kNnThreads, 0, stream>>>(dKeys, n, dim, metric, so3w, l2w, dQueries, nq, k, keysPerSlice,
                                                  dPartDist, dPartIdx);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    nnMergeKernel<<<(unsigned)((nq + 127) / 128), 128, 0, stream>>>(dPartDist, dPartIdx, nq, slices, k, dDist, dIdx);
    return cudaGetLastError();
}

}  // namespace mplb
