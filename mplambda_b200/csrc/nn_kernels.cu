// Batched exact nearest-neighbour kernels for sm_100a.
//
// Replaces the query side of nigh::Nigh as the planners use it (reference include/mpl/prrt.hpp:66-68
// `nn_.nearest(scale(q))`, include/mpl/pcforest.hpp:80-87 `nn_.nearest(nbh, scale(q), k)`): exact
// 1-NN / k-NN under nigh's SE3Space<S,a,b> metric  a*acos(|q.q'|) + b*||t-t'||  or L1Space<S,dim>
// ([EXT] nigh), ties resolved to the lowest insertion index, k-NN sorted ascending.
//
// Brute force, tiled: a CTA owns 128 queries (one per thread, key in registers) and one slice of
// the key set; key tiles are staged through shared memory with coalesced loads and every thread
// scans the tile (shared-memory broadcast reads).  Slices give the grid enough CTAs when there are
// few queries; merge kernels combine the per-slice lists.
//
// Exactness at FP32 speed: pass 1 scans an FP32 copy of the keys and finds, per query, the k-th
// smallest FP32 distance T.  With m bounding |d32 - d64| (acos amplifies the FP32 error of |q.q'|
// by at most sqrt(2 delta) near 1; coordinates round at 2^-24 relative), every true k-nearest key
// has d32 <= T + 2m, so pass 2 re-scans in FP32 and evaluates only those candidates in FP64 with
// the oracle's operation order (no FMA contraction): the L1 metric is bit-identical to the CPU
// path and the SE3 metric differs at most in acos' last bit; ties go to the lowest index.
#include "nn_device.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>

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

__device__ __forceinline__ float sqrtApprox(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// acos on [0,1] as sqrt(1-x) * cubic (Abramowitz & Stegun 4.4.45, |error| <= 6.8e-5): cheap, and its error
// is part of the candidate margin
__device__ __forceinline__ float acosApprox(float x) {
    const float p = fmaf(fmaf(fmaf(-0.0187293f, x, 0.0742610f), x, -0.2121144f), x, 1.5707288f);
    return sqrtApprox(fmaxf(1.f - x, 0.f)) * p;
}

// FP32 estimate of the distance, only if it can be <= bound: most keys are rejected by their translation
// alone (7 instructions) before any acos / sqrt.  -> true and d when d <= bound.
__device__ __forceinline__ bool nnWithin32(const float *q, const float *key, int dim, int metric, float so3w, float l2w,
                                           float bound, float &d) {
    if (metric == 0) {
        const float dx = key[4] - q[4], dy = key[5] - q[5], dz = key[6] - q[6];
        const float t2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (l2w * l2w * t2 > bound * bound) return false;
        const float x = fabsf(fmaf(key[3], q[3], fmaf(key[2], q[2], fmaf(key[1], q[1], key[0] * q[0]))));
        const float so3 = x < 1.f ? acosApprox(x) : 0.f;
        d = fmaf(so3w, so3, l2w * sqrtApprox(t2));
        return d <= bound;
    }
    float sum = 0.f;
    for (int i = 0; i < dim; ++i) sum += fabsf(key[i] - q[i]);
    d = sum;
    return d <= bound;
}

// bound of |d32 - d64| for one query (see the header comment); absMax = largest |coordinate| in play
__device__ __forceinline__ float nnMargin(int dim, int metric, float so3w, float l2w, float absMax) {
    if (metric == 0) return so3w * (2e-3f + 7e-5f) + l2w * 4e-6f * absMax + 1e-30f;
    return (float)dim * 1e-6f * absMax + 1e-30f;
}

// Pass 1.  grid = (query blocks, slices).  outVal: [nq][slices][k] ascending FP32 distances (+inf padded)
__global__ void __launch_bounds__(kNnThreads)
nnBound32Kernel(const float *__restrict__ keys32, unsigned long long n, int dim, int metric, float so3w, float l2w,
                const double *__restrict__ queries, unsigned long long nq, int k, unsigned long long keysPerSlice,
                unsigned long long stride, float *__restrict__ outVal) {
    // stride > 1: the scan runs over the sample {0, stride, 2 stride, ...} of the keys (n counts the sample)
    __shared__ __align__(16) float tile[kNnTile * kNnMaxDim];
    const unsigned long long qi = blockIdx.x * (unsigned long long)kNnThreads + threadIdx.x;
    const bool haveQuery = qi < nq;
    float q[kNnMaxDim];
#pragma unroll
    for (int i = 0; i < kNnMaxDim; ++i) q[i] = (haveQuery && i < dim) ? (float)__ldg(queries + qi * dim + i) : 0.f;
    float best[kNnMaxK];
    int cnt = 0;
    float b1 = INFINITY;
    const unsigned long long lo = blockIdx.y * keysPerSlice;
    const unsigned long long hi = min(n, lo + keysPerSlice);
    for (unsigned long long t0 = lo; t0 < hi; t0 += kNnTile) {
        const int tn = (int)min((unsigned long long)kNnTile, hi - t0);
        const int nd = tn * dim;
        __syncthreads();
        if (stride == 1) {
            for (int i = threadIdx.x; i < nd; i += kNnThreads) tile[i] = __ldg(keys32 + t0 * dim + i);
        } else {
            for (int i = threadIdx.x; i < nd; i += kNnThreads)
                tile[i] = __ldg(keys32 + (t0 + i / dim) * stride * dim + i % dim);
        }
        __syncthreads();
        if (!haveQuery) continue;
        if (k == 1) {
            for (int j = 0; j < tn; ++j) {
                float d;
                if (nnWithin32(q, tile + j * dim, dim, metric, so3w, l2w, b1, d)) b1 = d;
            }
        } else {
            for (int j = 0; j < tn; ++j) {
                float d;
                if (!nnWithin32(q, tile + j * dim, dim, metric, so3w, l2w, cnt == k ? best[k - 1] : INFINITY, d)) continue;
                if (cnt == k && !(d < best[k - 1])) continue;
                int pos = cnt < k ? cnt : k - 1;
                while (pos > 0 && best[pos - 1] > d) {
                    best[pos] = best[pos - 1];
                    --pos;
                }
                best[pos] = d;
                if (cnt < k) ++cnt;
            }
        }
    }
    if (!haveQuery) return;
    float *ov = outVal + (qi * gridDim.y + blockIdx.y) * k;
    if (k == 1) ov[0] = b1;
    else
        for (int r = 0; r < k; ++r) ov[r] = r < cnt ? best[r] : INFINITY;
}

// k > 1, pass 1b.  grid = (query blocks, slices).  Sorted insertion per key is what made the k-NN bound pass 17x
// slower than the 1-NN one (a warp pays for an insertion whenever any of its lanes inserts), and the pass only has
// to produce a threshold with at least k keys under it.  So: the k-th smallest distance T0 over a strided SAMPLE
// of the keys (nnBound32Kernel, cheap) is such a threshold -- with about k n / m keys under it -- and this kernel
// COUNTS, over all keys, how many lie under each of kNnLevels fractions of T0 (registers only); the smallest
// level that still holds k keys becomes the threshold of pass 2.
#ifndef MPLB_NN_LEVELS
#define MPLB_NN_LEVELS 8
#endif
constexpr int kNnLevels = MPLB_NN_LEVELS;
constexpr int kNnCandCap = 128;   // candidates a thread notes before it works them off (more simply means another round)

__global__ void __launch_bounds__(kNnThreads)
nnCountKernel(const float *__restrict__ keys32, unsigned long long n, int dim, int metric, float so3w, float l2w,
              const double *__restrict__ queries, unsigned long long nq, int k, unsigned long long keysPerSlice,
              const float *__restrict__ sampleVal, float ratio, unsigned *__restrict__ counts) {
    __shared__ __align__(16) float tile[kNnTile * kNnMaxDim];
    const unsigned long long qi = blockIdx.x * (unsigned long long)kNnThreads + threadIdx.x;
    const bool haveQuery = qi < nq;
    float q[kNnMaxDim];
#pragma unroll
    for (int i = 0; i < kNnMaxDim; ++i) q[i] = (haveQuery && i < dim) ? (float)__ldg(queries + qi * dim + i) : 0.f;
    const float t0v = haveQuery ? sampleVal[qi * k + (k - 1)] : -1.f;   // +inf: the sample held fewer than k keys
    float lev[kNnLevels];
    unsigned cnt[kNnLevels];
    {
        float f = 1.f;
#pragma unroll
        for (int j = kNnLevels - 1; j >= 0; --j) {
            lev[j] = t0v * f;
            cnt[j] = 0;
            f *= ratio;
        }
    }
    const unsigned long long lo = blockIdx.y * keysPerSlice;
    const unsigned long long hi = min(n, lo + keysPerSlice);
    for (unsigned long long t0 = lo; t0 < hi; t0 += kNnTile) {
        const int tn = (int)min((unsigned long long)kNnTile, hi - t0);
        const int nd = tn * dim;
        __syncthreads();
        for (int i = threadIdx.x; i < nd; i += kNnThreads) tile[i] = __ldg(keys32 + t0 * dim + i);
        __syncthreads();
        if (!haveQuery || !(t0v < INFINITY)) continue;
        for (int j = 0; j < tn; ++j) {
            float d;
            if (!nnWithin32(q, tile + j * dim, dim, metric, so3w, l2w, t0v, d)) continue;
#pragma unroll
            for (int l = 0; l < kNnLevels; ++l) cnt[l] += d <= lev[l] ? 1u : 0u;
        }
    }
    if (!haveQuery) return;
#pragma unroll
    for (int l = 0; l < kNnLevels; ++l)
        if (cnt[l]) atomicAdd(counts + qi * kNnLevels + l, cnt[l]);
}

// k > 1: the smallest counted level that holds at least k keys (the last level, T0 itself, always does)
__global__ void nnLevelKernel(const float *__restrict__ sampleVal, const unsigned *__restrict__ counts, float ratio,
                              unsigned long long nq, int k, float *__restrict__ kth) {
    const unsigned long long qi = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const float t0v = sampleVal[qi * k + (k - 1)];
    float best = t0v;
    if (t0v < INFINITY) {
        float f = 1.f;
        for (int j = kNnLevels - 1; j >= 0; --j) {   // same products as nnCountKernel
            if (counts[qi * kNnLevels + j] >= (unsigned)k) best = t0v * f;
            f *= ratio;
        }
    }
    kth[qi] = best;
}

// per query: k-th smallest FP32 distance over all slices, plus twice the error margin
__global__ void nnThresholdKernel(const float *__restrict__ partVal, const double *__restrict__ queries,
                                  unsigned long long nq, int slices, int k, int dim, int metric, float so3w, float l2w,
                                  float keysAbsMax, const float *__restrict__ kthIn, float *__restrict__ thr) {
    const unsigned long long qi = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const float *v = partVal + qi * slices * k;
    unsigned char head[256];
    for (int s = 0; s < slices; ++s) head[s] = 0;
    float kth = INFINITY;
    if (kthIn) kth = kthIn[qi];   // k > 1: nnLevelKernel already chose a level with at least k keys under it
    for (int r = 0; r < k && !kthIn; ++r) {
        int bs = -1;
        float bd = INFINITY;
        for (int s = 0; s < slices; ++s)
            if (head[s] < k && v[s * k + head[s]] < bd) {
                bd = v[s * k + head[s]];
                bs = s;
            }
        if (bs < 0) {   // fewer than k keys: everything is a candidate
            kth = INFINITY;
            break;
        }
        ++head[bs];
        kth = bd;
    }
    float absMax = keysAbsMax;
    if (metric == 0) {   // the FP32 bound assumes unit quaternions; anything else is handled exactly
        float qn = 0.f;
        for (int i = 0; i < 4; ++i) qn += (float)queries[qi * dim + i] * (float)queries[qi * dim + i];
        if (!(fabsf(qn - 1.f) < 1e-3f)) absMax = INFINITY;
    }
    for (int i = (metric == 0 ? 4 : 0); i < dim; ++i) absMax = fmaxf(absMax, fabsf((float)queries[qi * dim + i]) * 1.0000002f);
    thr[qi] = kth + 2.f * nnMargin(dim, metric, so3w, l2w, absMax) + 8e-6f * fabsf(kth);
}

// Pass 2.  grid = (query blocks, slices).  FP32 re-scan; keys within the query's threshold are evaluated in
// FP64.  out*: [nq][slices][k]
__global__ void __launch_bounds__(kNnThreads)
nnRefineKernel(const float *__restrict__ keys32, const double *__restrict__ keys, unsigned long long n, int dim, int metric,
               double so3w, double l2w, const double *__restrict__ queries, const float *__restrict__ thr,
               unsigned long long nq, int k, unsigned long long keysPerSlice, double *__restrict__ outDist,
               long long *__restrict__ outIdx, unsigned long long *exactCount) {
    __shared__ __align__(16) float tile[kNnTile * kNnMaxDim];
    const unsigned long long qi = blockIdx.x * (unsigned long long)kNnThreads + threadIdx.x;
    const bool haveQuery = qi < nq;
    double q[kNnMaxDim];
    float q32[kNnMaxDim];
#pragma unroll
    for (int i = 0; i < kNnMaxDim; ++i) {
        q[i] = (haveQuery && i < dim) ? __ldg(queries + qi * dim + i) : 0.0;
        q32[i] = (float)q[i];
    }
    const float limit = haveQuery ? thr[qi] : -1.f;
    const float so3w32 = (float)so3w, l2w32 = (float)l2w;
    double bestD[kNnMaxK];
    long long bestI[kNnMaxK];
    int cnt = 0;
    double d1 = INFINITY;
    long long i1 = -1;
    unsigned exact = 0;
    const unsigned long long lo = blockIdx.y * keysPerSlice;
    const unsigned long long hi = min(n, lo + keysPerSlice);
    unsigned cand[kNnCandCap];   // candidate keys of this slice, in index order (offsets from lo)
    int nc = 0;
    auto flush = [&]() {
        for (int c = 0; c < nc; ++c) {
            const unsigned long long ki = lo + cand[c];
            double key[kNnMaxDim];
            const double *kp = keys + ki * dim;
            for (int i = 0; i < dim; ++i) key[i] = __ldg(kp + i);
            const double d = nnDistance(q, key, dim, metric, so3w, l2w);
            ++exact;
            if (cnt == k && !(d < bestD[k - 1])) continue;
            int pos = cnt < k ? cnt : k - 1;
            while (pos > 0 && bestD[pos - 1] > d) {   // equal distances keep their (index) order
                bestD[pos] = bestD[pos - 1];
                bestI[pos] = bestI[pos - 1];
                --pos;
            }
            bestD[pos] = d;
            bestI[pos] = (long long)ki;
            if (cnt < k) ++cnt;
        }
        nc = 0;
    };
    for (unsigned long long t0 = lo; t0 < hi; t0 += kNnTile) {
        const int tn = (int)min((unsigned long long)kNnTile, hi - t0);
        const int nd = tn * dim;
        __syncthreads();
        for (int i = threadIdx.x; i < nd; i += kNnThreads) tile[i] = __ldg(keys32 + t0 * dim + i);
        __syncthreads();
        if (!haveQuery) continue;
        for (int j = 0; j < tn; ++j) {
            float d32;
            if (!nnWithin32(q32, tile + j * dim, dim, metric, so3w32, l2w32, limit, d32)) continue;
            if (k == 1) {
                double key[kNnMaxDim];
                const double *kp = keys + (t0 + j) * dim;
                for (int i = 0; i < dim; ++i) key[i] = __ldg(kp + i);
                const double d = nnDistance(q, key, dim, metric, so3w, l2w);
                ++exact;
                if (d < d1) {
                    d1 = d;
                    i1 = (long long)(t0 + j);
                }
                continue;
            }
            // k > 1: only note the candidate here (one store); the FP64 distances and the sorted insertions run
            // once the scan is over, when every lane of the warp has its list to work through -- a lane that
            // evaluates and inserts in the middle of the scan makes the other 31 wait
            cand[nc++] = (unsigned)(t0 + j - lo);
            if (nc == kNnCandCap) flush();
        }
    }
    if (haveQuery && k > 1) flush();
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
    if (exactCount && exact) atomicAdd(exactCount, (unsigned long long)exact);
}

__global__ void nnToFloatKernel(const double *__restrict__ src, float *__restrict__ dst, unsigned long long n) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)src[i];
}

// one thread per query: merge `slices` ascending lists of length k into one (ties -> lower index)
// {largest |translation coordinate| (SE3) / |coordinate| (L1), 1 if an SE3 key's quaternion is not unit} of keys that
// arrived in device memory; non-negative floats order like their bit patterns (a NaN or an infinity ends up on top)
__global__ void nnKeyStatsKernel(const double *__restrict__ keys, unsigned long long n, int dim, int metric,
                                 unsigned *__restrict__ out) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    float m = 0.f;
    unsigned nonUnit = 0;
    if (i < n) {
        const double *q = keys + i * dim;
        for (int d = (metric == 0 ? 4 : 0); d < dim; ++d) {
            const float v = __double2float_ru(fabs(q[d]));
            m = (v > m || v != v) ? v : m;
        }
        if (metric == 0) {
            const double qn = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
            nonUnit = !(fabs(qn - 1.0) < 1e-3);
        }
    }
    unsigned bits = __float_as_uint(m);
    for (int o = 16; o; o >>= 1) {
        bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
        nonUnit |= __shfl_xor_sync(0xffffffffu, nonUnit, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, bits);
        if (nonUnit) atomicExch(out + 1, __float_as_uint(1.f));
    }
}

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

cudaError_t launchNnToFloat(const double *dSrc, float *dDst, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    nnToFloatKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dSrc, dDst, n);
    return cudaGetLastError();
}

cudaError_t launchNnKeyStats(const double *dKeys, size_t n, int dim, int metric, float *dOut, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(dOut, 0, 2 * sizeof(float), stream);
    if (e != cudaSuccess || n == 0) return e;
    nnKeyStatsKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dKeys, n, dim, metric, (unsigned *)dOut);
    return cudaGetLastError();
}

cudaError_t launchNnSearch(const double *dKeys, const float *dKeys32, float keysAbsMax, size_t n, int dim, int metric,
                           double so3w, double l2w, const double *dQueries, size_t nq, int k, int slices,
                           float *dPartVal, float *dThr, double *dPartDist, long long *dPartIdx, double *dDist,
                           long long *dIdx, unsigned long long *dExactCount, cudaStream_t stream) {
    if (nq == 0) return cudaSuccess;
    const size_t keysPerSlice = std::max<size_t>(((n + slices - 1) / slices + kNnTile - 1) / kNnTile * kNnTile, kNnTile);
    dim3 grid((unsigned)((nq + kNnThreads - 1) / kNnThreads), (unsigned)slices);
    if (k == 1) {
        nnBound32Kernel<<<grid, kNnThreads, 0, stream>>>(dKeys32, n, dim, metric, (float)so3w, (float)l2w, dQueries, nq, k,
                                                         keysPerSlice, 1, dPartVal);
        nnThresholdKernel<<<(unsigned)((nq + 127) / 128), 128, 0, stream>>>(dPartVal, dQueries, nq, slices, k, dim, metric,
                                                                            (float)so3w, (float)l2w, keysAbsMax, nullptr, dThr);
    } else {
        // sample of ~16 k keys (the sorted insertions of the sample scan cost more than its distances: measured 2.5 ms
        // for 3 277 sampled keys against 0.5 ms for counting over all 16 384); dPartVal: [nq][k] sample distances,
        // then [nq][kNnLevels] counts, then [nq] chosen levels
        const size_t want = (size_t)16 * k;
        const size_t stride = std::max<size_t>(1, n / std::max<size_t>(want, 1));
        const size_t m = (n + stride - 1) / stride;
        const size_t sampleSlice = std::max<size_t>((m + kNnTile - 1) / kNnTile * kNnTile, kNnTile);
        // levels: k ... k n / m keys expected under them, geometric; distances scale with count^(1/dimension)
        const double R = std::max(1.0, (double)n / (double)m);
        const float ratio = (float)std::pow(1.0 / R, 1.0 / ((kNnLevels - 1) * (metric == 0 ? 6.0 : (double)dim)));
        unsigned *dCounts = (unsigned *)(dPartVal + nq * k);
        float *dKth = (float *)(dCounts + nq * kNnLevels);
        cudaError_t e0 = cudaMemsetAsync(dCounts, 0, nq * kNnLevels * sizeof(unsigned), stream);
        if (e0 != cudaSuccess) return e0;
        nnBound32Kernel<<<dim3(grid.x, 1), kNnThreads, 0, stream>>>(dKeys32, m, dim, metric, (float)so3w, (float)l2w, dQueries,
                                                                    nq, k, sampleSlice, stride, dPartVal);
        nnCountKernel<<<grid, kNnThreads, 0, stream>>>(dKeys32, n, dim, metric, (float)so3w, (float)l2w, dQueries, nq, k,
                                                       keysPerSlice, dPartVal, ratio, dCounts);
        nnLevelKernel<<<(unsigned)((nq + 127) / 128), 128, 0, stream>>>(dPartVal, dCounts, ratio, nq, k, dKth);
        nnThresholdKernel<<<(unsigned)((nq + 127) / 128), 128, 0, stream>>>(dPartVal, dQueries, nq, 1, k, dim, metric,
                                                                            (float)so3w, (float)l2w, keysAbsMax, dKth, dThr);
    }
    if (slices == 1) {
        nnRefineKernel<<<grid, kNnThreads, 0, stream>>>(dKeys32, dKeys, n, dim, metric, so3w, l2w, dQueries, dThr, nq, k,
                                                        keysPerSlice, dDist, dIdx, dExactCount);
        return cudaGetLastError();
    }
    nnRefineKernel<<<grid, kNnThreads, 0, stream>>>(dKeys32, dKeys, n, dim, metric, so3w, l2w, dQueries, dThr, nq, k,
                                                    keysPerSlice, dPartDist, dPartIdx, dExactCount);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    nnMergeKernel<<<(unsigned)((nq + 127) / 128), 128, 0, stream>>>(dPartDist, dPartIdx, nq, slices, k, dDist, dIdx);
    return cudaGetLastError();
}

}  // namespace mplb
