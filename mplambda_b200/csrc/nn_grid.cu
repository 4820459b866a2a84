// Insert-friendly spatial index for the SE(3) nearest-neighbour structure (sm_100a).
//
// Replaces what nigh's KD-tree strategy gives the planners beyond brute force (reference include/mpl/prrt.hpp:35-42,
// 61-68 -- Nigh<Node*, Space, NodeKey, Concurrent, auto_strategy_t<Space, Concurrent>> -- and include/mpl/pcforest.hpp:
// 42-51, 80-87): exact 1-NN / k-NN in o(n) per query while nodes keep being inserted.
//
// nigh's SE3Space distance is  so3w * acos|q.q'| + l2w * ||t - t'||  >=  l2w * ||t - t'||, so a uniform grid over the
// TRANSLATIONS prunes exactly: keys are counting-sorted by cell (cells x-fastest, so a row of cells is one contiguous
// run of keys), a query is answered by one warp that scans the cube of cells around its own cell ring by ring:
//   pass 1 (FP32) histograms the distances it meets over geometric levels (2^(1/6) apart: a factor 2 in the number of
//          keys of a 6-dimensional ball) and stops at the first ring r and level T with at least k keys under T and
//          T + 2m below the distance to anything outside the cube (m bounds |d32 - d64|, as in nn_kernels.cu);
//   pass 2 rescans the cube, evaluates keys with d32 <= T + 2m in FP64 with the oracle's operation order and ranks them
//          by (distance, insertion index): the same results, bit for bit, as the brute-force kernels.
// Too many candidates for the warp's shared-memory list (duplicate keys) fall back to k rounds of minimum extraction
// over the same cube -- exact, only slower.  The structure is rebuilt (three kernels and a scan, O(n)) when the keys
// inserted since the last build exceed a quarter of it; until then they are searched by the brute-force kernels and
// the two sorted lists are merged (nn_capi.cu), which keeps insertion O(1) amortised -- the log-structured pattern.
// L1Space<S, dim> keys (the Fetch planners' joint space, fetch_scenario.hpp:19,42-45) are indexed the same way over
// their first three coordinates: sum_i |x_i - y_i| >= the distance within any subset of the coordinates.
#include "nn_device.cuh"

#include <cfloat>
#include <cmath>
#include <cstring>

namespace mplb {
namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kGridWarps = 8;          // queries (warps) per CTA
constexpr int kGridBins = 64;          // histogram levels per query
constexpr int kGridCand = 160;         // shared-memory candidates per query (k <= 64)
constexpr int kGridMaxRing = 12;       // beyond that the cube becomes the whole grid

// order-preserving float <-> unsigned (for atomicMin / atomicMax on floats of either sign)
__device__ __forceinline__ unsigned orderedBits(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void nnGridBoundsKernel(const double *__restrict__ keys, unsigned n, int dim, int coord0, unsigned *__restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned lo[3] = {~0u, ~0u, ~0u}, hi[3] = {0u, 0u, 0u};
    if (i < n) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v = keys[(unsigned long long)dim * i + coord0 + a];
            lo[a] = orderedBits(__double2float_rd(v));
            hi[a] = orderedBits(__double2float_ru(v));
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(kFullMask, lo[a]);
        hi[a] = __reduce_max_sync(kFullMask, hi[a]);
    }
    if ((threadIdx.x & 31u) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(out + a, lo[a]);
            atomicMax(out + 3 + a, hi[a]);
        }
    }
}

__device__ __forceinline__ int cellCoord(double v, double origin, double invCell, int dim) {
    const double c = floor((v - origin) * invCell);
    return c < 0.0 ? 0 : c >= (double)dim ? dim - 1 : (int)c;
}

__global__ void nnGridCountKernel(const double *__restrict__ keys, unsigned n, NnGridGeom g, unsigned *__restrict__ cellOf,
                                  unsigned *__restrict__ counts) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *k = keys + (unsigned long long)g.dim * i + g.coord0;
    const int x = cellCoord(k[0], g.origin[0], g.invCell, g.dims[0]);
    const int y = cellCoord(k[1], g.origin[1], g.invCell, g.dims[1]);
    const int z = cellCoord(k[2], g.origin[2], g.invCell, g.dims[2]);
    const unsigned c = ((unsigned)z * g.dims[1] + y) * g.dims[0] + x;
    cellOf[i] = c;
    atomicAdd(counts + c, 1u);
}

// exclusive prefix sum of counts[0, n) into start[0, n], one CTA: every thread sums a contiguous run, the run totals
// are scanned in shared memory, every thread writes its run's prefixes
__global__ void __launch_bounds__(1024) nnGridScanKernel(const unsigned *__restrict__ counts, unsigned n, unsigned *__restrict__ start) {
    __shared__ unsigned totals[1024];
    const unsigned per = (n + 1023u) / 1024u, lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
    unsigned s = 0;
    for (unsigned i = lo; i < hi; ++i) s += counts[i];
    totals[threadIdx.x] = s;
    __syncthreads();
    for (unsigned d = 1; d < 1024; d <<= 1) {
        const unsigned v = threadIdx.x >= d ? totals[threadIdx.x - d] : 0u;
        __syncthreads();
        totals[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = totals[threadIdx.x] - s;
    for (unsigned i = lo; i < hi; ++i) {
        start[i] = run;
        run += counts[i];
    }
    if (threadIdx.x == 1023) start[n] = totals[1023];
}

__global__ void nnGridScatterKernel(const double *__restrict__ keys, unsigned n, int dim, const unsigned *__restrict__ cellOf,
                                    const unsigned *__restrict__ start, unsigned *__restrict__ fill,
                                    unsigned *__restrict__ perm, float4 *__restrict__ sorted) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned c = cellOf[i];
    const unsigned pos = start[c] + atomicAdd(fill + c, 1u);
    perm[pos] = i;
    const double *k = keys + (unsigned long long)dim * i;
    float f[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) f[d] = d < dim ? (float)k[d] : 0.f;
    sorted[2ull * pos] = make_float4(f[0], f[1], f[2], f[3]);
    sorted[2ull * pos + 1] = make_float4(f[4], f[5], f[6], f[7]);
}

__device__ __forceinline__ float sqrtApproxG(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// the FP32 estimate of nn_kernels.cu (same polynomial acos, same margin) / the L1 sum
template <int kMetric>
__device__ __forceinline__ float gridDist32(const float q[8], const float4 ka, const float4 kb, float so3w, float l2w) {
    if (kMetric == 0) {
        const float dx = kb.x - q[4], dy = kb.y - q[5], dz = kb.z - q[6];
        const float t2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const float x = fabsf(fmaf(ka.w, q[3], fmaf(ka.z, q[2], fmaf(ka.y, q[1], ka.x * q[0]))));
        const float p = fmaf(fmaf(fmaf(-0.0187293f, x, 0.0742610f), x, -0.2121144f), x, 1.5707288f);
        const float so3 = x < 1.f ? sqrtApproxG(fmaxf(1.f - x, 0.f)) * p : 0.f;
        return fmaf(so3w, so3, l2w * sqrtApproxG(t2));
    }
    // (coordinates past the key's dimension are zero in both the sorted copy and q)
    return ((fabsf(ka.x - q[0]) + fabsf(ka.y - q[1])) + (fabsf(ka.z - q[2]) + fabsf(ka.w - q[3]))) +
           ((fabsf(kb.x - q[4]) + fabsf(kb.y - q[5])) + (fabsf(kb.z - q[6]) + fabsf(kb.w - q[7])));
}
template <int kMetric>
__device__ __forceinline__ double gridDist64(const double q[8], const double *key, int dim, double so3w, double l2w) {
    if (kMetric == 0) {
        const double d = fabs(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(key[0], q[0]), __dmul_rn(key[1], q[1])),
                                                  __dmul_rn(key[2], q[2])),
                                        __dmul_rn(key[3], q[3])));
        const double so3 = d < 1.0 ? acos(d) : 0.0;
        const double dx = __dsub_rn(key[4], q[4]), dy = __dsub_rn(key[5], q[5]), dz = __dsub_rn(key[6], q[6]);
        const double l2 = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        return __dadd_rn(__dmul_rn(so3w, so3), __dmul_rn(l2w, l2));
    }
    double s = 0;   // nn_kernels.cu nnDistance: the oracle's order
    for (int i = 0; i < dim; ++i) s = __dadd_rn(s, fabs(__dsub_rn(key[i], q[i])));
    return s;
}

struct Cube {
    int lo[3], hi[3];   // inclusive cell ranges
};

// Visits the keys of the cells of `cube` that are not in `inner` (inner.lo[0] > inner.hi[0]: nothing excluded), a row
// of cells (contiguous keys) at a time, lanes striding over the keys: f(position in the sorted arrays, position valid);
// every lane of the warp makes every call.
template <typename F>
__device__ __forceinline__ void forKeys(const NnGridGeom &g, const unsigned *__restrict__ start, const Cube &cube,
                                        const Cube &inner, unsigned lane, F f) {
    const bool haveInner = inner.lo[0] <= inner.hi[0];
    for (int z = cube.lo[2]; z <= cube.hi[2]; ++z)
        for (int y = cube.lo[1]; y <= cube.hi[1]; ++y) {
            const unsigned row = ((unsigned)z * g.dims[1] + y) * g.dims[0];
            const bool split = haveInner && z >= inner.lo[2] && z <= inner.hi[2] && y >= inner.lo[1] && y <= inner.hi[1];
            // the part of the row left of the inner cube, then the part right of it (or the whole row)
            for (int part = 0; part < (split ? 2 : 1); ++part) {
                const int x0 = split ? (part == 0 ? cube.lo[0] : inner.hi[0] + 1) : cube.lo[0];
                const int x1 = split ? (part == 0 ? inner.lo[0] - 1 : cube.hi[0]) : cube.hi[0];
                if (x0 > x1) continue;
                const unsigned b = __ldg(start + row + x0), e = __ldg(start + row + x1 + 1);
                for (unsigned at = b; at < e; at += 32) f(at + lane, at + lane < e);   // (warp-uniform trip count)
            }
        }
}

// one warp per query
template <int kMetric>
__global__ void __launch_bounds__(kGridWarps * 32)
nnGridSearchKernel(NnGridGeom g, const unsigned *__restrict__ start, const unsigned *__restrict__ perm,
                   const float4 *__restrict__ sorted, const double *__restrict__ keys, unsigned nBuilt, double so3w,
                   double l2w, float keysAbsMax, const double *__restrict__ queries, unsigned long long nq, int k,
                   double *__restrict__ outDist, long long *__restrict__ outIdx, unsigned long long outStride,
                   unsigned long long *__restrict__ stats) {
    __shared__ unsigned hist[kGridWarps][kGridBins];
    __shared__ double candD[kGridWarps][kGridCand];
    __shared__ unsigned candI[kGridWarps][kGridCand];
    __shared__ unsigned candN[kGridWarps];
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const unsigned long long qi = blockIdx.x * (unsigned long long)kGridWarps + w;
    if (qi >= nq) return;
    const int dim = g.dim, c0 = g.coord0;
    double q[8];
    float q32[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        q[i] = i < dim ? __ldg(queries + qi * dim + i) : 0.0;
        q32[i] = (float)q[i];
    }
    const float so3w32 = (float)so3w, l2w32 = kMetric == 0 ? (float)l2w : 1.f;   // (L1: the grid coordinates weigh 1)
    // |d32 - d64| <= m (nn_kernels.cu: nnMargin); a query whose quaternion is not unit, or that is not finite, gets no
    // FP32 filtering at all (everything in reach is evaluated exactly)
    float absMax = keysAbsMax;
#pragma unroll
    for (int i = (kMetric == 0 ? 4 : 0); i < (kMetric == 0 ? 7 : 8); ++i) absMax = fmaxf(absMax, fabsf(q32[i]) * 1.0000002f);
    const float qn = q32[0] * q32[0] + q32[1] * q32[1] + q32[2] * q32[2] + q32[3] * q32[3];
    const bool exactOnly = (kMetric == 0 && !(fabsf(qn - 1.f) < 1e-3f)) || !(absMax < INFINITY);
    const float m = kMetric == 0 ? so3w32 * (2e-3f + 7e-5f) + l2w32 * 4e-6f * absMax + 1e-30f
                                 : (float)dim * 1e-6f * absMax + 1e-30f;   // nn_kernels.cu: nnMargin
    // the three coordinates the grid is laid over
    double qg[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) qg[a] = kMetric == 0 ? q[4 + a] : q[a];
    (void)c0;
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) c[a] = cellCoord(qg[a], g.origin[a], g.invCell, g.dims[a]);
    // histogram levels: level j holds distances <= d0 * 2^(j / 6)
    // (the top level covers every key there is: the full rotation range plus the way to the farthest corner of the
    // grid, so that a query far outside the keys' box -- whose cube ends up being the whole grid -- still gets a finite
    // threshold; 64 levels span a factor 2^(63/6) = 1448 below that)
    float far2 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = fabsf((float)(qg[a] - g.origin[a])), hi = fabsf((float)(g.origin[a] + g.dims[a] * g.cell - qg[a]));
        far2 = kMetric == 0 ? fmaf(fmaxf(lo, hi), fmaxf(lo, hi), far2) : far2 + fmaxf(lo, hi);
    }
    float dTop;
    if (kMetric == 0) {
        dTop = (so3w32 * 1.5708f + l2w32 * sqrtf(far2)) * 1.01f;
    } else {   // the grid coordinates' share plus, for every other coordinate, the farthest a key can be
        dTop = far2;
        for (int i = 3; i < dim; ++i) dTop += fabsf(q32[i]) + (keysAbsMax < INFINITY ? keysAbsMax : 0.f);
        dTop *= 1.01f;
    }
    const float d0 = fmaxf(dTop * (1.f / 1448.f), 1e-20f);
    const float invD0 = 1.f / d0;
    for (int j = lane; j < kGridBins; j += 32) hist[w][j] = 0u;
    if (lane == 0) candN[w] = 0u;
    __syncwarp();

    Cube cube, inner;
    inner.lo[0] = 1; inner.hi[0] = 0;   // nothing scanned yet
    float T = INFINITY, Tlo = 0.f;      // candidate threshold (FP32 distance); (Tlo, T] is the level it came from
    unsigned under = 0;                 // keys counted under T
    bool whole = false;                 // the cube is the whole grid
    for (int r = 0; !exactOnly; ++r) {
        whole = true;
        float reach = INFINITY;   // distance from the query to the nearest face of the cube with cells beyond it
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            cube.lo[a] = max(c[a] - r, 0);
            cube.hi[a] = min(c[a] + r, g.dims[a] - 1);
            if (r > kGridMaxRing) {
                cube.lo[a] = 0;
                cube.hi[a] = g.dims[a] - 1;
            }
            if (cube.lo[a] > 0) {
                whole = false;
                reach = fminf(reach, (float)(qg[a] - (g.origin[a] + cube.lo[a] * g.cell)));
            }
            if (cube.hi[a] < g.dims[a] - 1) {
                whole = false;
                reach = fminf(reach, (float)((g.origin[a] + (cube.hi[a] + 1) * g.cell) - qg[a]));
            }
        }
        forKeys(g, start, cube, inner, lane, [&](unsigned p, bool have) {
            if (!have) return;
            const float4 ka = __ldg(sorted + 2ull * p), kb = __ldg(sorted + 2ull * p + 1);
            const float d = gridDist32<kMetric>(q32, ka, kb, so3w32, l2w32);
            // level of d: smallest j with d <= d0 * 2^(j/6)
            int j = d <= d0 ? 0 : (int)ceilf(6.f * __log2f(d * invD0) * 1.000001f);
            if (j < kGridBins && d == d) atomicAdd(&hist[w][j], 1u);
        });
        __syncwarp();
        inner = cube;
        // smallest level with at least k keys under it
        const unsigned h0 = hist[w][2 * lane], h1 = hist[w][2 * lane + 1];
        unsigned incl = h0 + h1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(kFullMask, incl, d);
            if ((int)lane >= d) incl += v;
        }
        const unsigned before = incl - h0 - h1;
        int lvl = before + h0 >= (unsigned)k ? 2 * (int)lane : before + h0 + h1 >= (unsigned)k ? 2 * (int)lane + 1 : 1 << 20;
        lvl = __reduce_min_sync(kFullMask, lvl);
        if (lvl < kGridBins) {
            // level bound, rounded up (exp2f is within 2 ulp)
            const float t = d0 * exp2f((float)lvl * (1.f / 6.f)) * 1.00001f;
            // everything outside the cube is at least l2w * reach away (FP64 geometry; a cell of slack against the
            // rounding of the cell assignment is far more than needed and costs nothing)
            if (whole || t + 2.f * m + 8e-6f * t < l2w32 * reach * 0.999999f) {
                T = t;
                Tlo = lvl > 0 ? d0 * exp2f((float)(lvl - 1) * (1.f / 6.f)) * 0.9999f : 0.f;
                // keys under T: all of the levels up to lvl (the last one is what a finer threshold could shed)
                under = __shfl_sync(kFullMask, before + h0 + ((lvl & 1) ? h1 : 0u), lvl >> 1);
                break;
            }
        }
        if (whole) break;   // fewer than k keys under any level: T stays +inf, everything is a candidate
    }
    if (exactOnly) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            cube.lo[a] = 0;
            cube.hi[a] = g.dims[a] - 1;
        }
    }
    // A level is 12 % wide.  Around a query inside the keys' cloud that is a factor two in keys; for a query far from it
    // one level can hold thousands (a thin shell of a large sphere cuts a whole slab out of the cloud).  While far more
    // keys than k lie under T the level (Tlo, T] is cut into 64 equal parts and the part holding the k-th key becomes
    // the new level: every key under T is inside the cube, so a pass over the cube sees them all.
    for (int refine = 0; refine < 3 && T < INFINITY && under > (unsigned)k + kGridCand / 3u; ++refine) {
        for (int j = lane; j < kGridBins; j += 32) hist[w][j] = 0u;
        __syncwarp();
        const float scale = (float)kGridBins / (T - Tlo);
        unsigned belowLane = 0;
        inner.lo[0] = 1; inner.hi[0] = 0;
        forKeys(g, start, cube, inner, lane, [&](unsigned p, bool have) {
            if (!have) return;
            const float4 ka = __ldg(sorted + 2ull * p), kb = __ldg(sorted + 2ull * p + 1);
            const float d = gridDist32<kMetric>(q32, ka, kb, so3w32, l2w32);
            if (d <= Tlo) ++belowLane;
            else if (d <= T) atomicAdd(&hist[w][min(kGridBins - 1, (int)((d - Tlo) * scale))], 1u);
        });
        __syncwarp();
        const unsigned below = __reduce_add_sync(kFullMask, belowLane);
        const unsigned h0 = hist[w][2 * lane], h1 = hist[w][2 * lane + 1];
        unsigned incl = h0 + h1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(kFullMask, incl, d);
            if ((int)lane >= d) incl += v;
        }
        const unsigned before = below + incl - h0 - h1;
        int part = before + h0 >= (unsigned)k ? 2 * (int)lane : before + h0 + h1 >= (unsigned)k ? 2 * (int)lane + 1 : 1 << 20;
        part = __reduce_min_sync(kFullMask, part);
        if (part >= kGridBins) break;   // (cannot happen: at least k keys lie under T) keep the level as it is
        if (below >= (unsigned)k) {     // (rounding at the lower edge) the k-th key is not above Tlo after all
            T = Tlo * 1.0001f;
            break;
        }
        const float width = (T - Tlo) * (1.f / (float)kGridBins);
        const float newLo = Tlo + width * (float)part * 0.99999f;
        const float newT = fminf(T, (Tlo + width * (float)(part + 1)) * 1.00001f);
        under = __shfl_sync(kFullMask, before + h0 + ((part & 1) ? h1 : 0u), part >> 1);
        Tlo = fminf(newLo, newT);
        T = newT;
        if (!(T > Tlo)) break;
    }
    const float limit = exactOnly ? INFINITY : T + 2.f * m + 8e-6f * fabsf(T);
    inner.lo[0] = 1; inner.hi[0] = 0;
    // pass 2: exact distances of the candidates
    unsigned exact = 0;
    forKeys(g, start, cube, inner, lane, [&](unsigned p, bool have) {
        bool cand = have;
        if (have && !exactOnly) {
            const float4 ka = __ldg(sorted + 2ull * p), kb = __ldg(sorted + 2ull * p + 1);
            cand = gridDist32<kMetric>(q32, ka, kb, so3w32, l2w32) <= limit;
        }
        const unsigned mk = __ballot_sync(kFullMask, cand);
        if (!mk) return;
        unsigned base = 0;
        const int leader = __ffs(mk) - 1;
        if ((int)lane == leader) base = atomicAdd(&candN[w], (unsigned)__popc(mk));
        base = __shfl_sync(kFullMask, base, leader);
        if (cand) {
            const unsigned at = base + __popc(mk & ((1u << lane) - 1u));
            if (at < (unsigned)kGridCand) {
                const unsigned ki = __ldg(perm + p);
                double key[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) key[i] = i < dim ? __ldg(keys + (unsigned long long)dim * ki + i) : 0.0;
                candD[w][at] = gridDist64<kMetric>(q, key, dim, so3w, l2w);
                candI[w][at] = ki;
                ++exact;
            }
        }
    });
    __syncwarp();
    const unsigned nc = candN[w];
    double *od = outDist + qi * outStride;
    long long *oi = outIdx + qi * outStride;
    if (nc <= (unsigned)kGridCand) {
        // rank by (distance, index)
        for (unsigned a = lane; a < nc; a += 32) {
            const double d = candD[w][a];
            const unsigned ix = candI[w][a];
            unsigned rank = 0;
            for (unsigned b = 0; b < nc; ++b) {
                const double e = candD[w][b];
                rank += (e < d || (e == d && candI[w][b] < ix)) ? 1u : 0u;
            }
            if (rank < (unsigned)k) {
                od[rank] = d;
                oi[rank] = (long long)ix;
            }
        }
        for (unsigned r = nc + lane; r < (unsigned)k; r += 32) {
            od[r] = INFINITY;
            oi[r] = -1;
        }
    } else {
        // more candidates than the list holds (duplicate keys): k rounds of minimum extraction over the cube, each
        // picking the smallest (distance, index) above the previous pick -- exact, slow, rare
        double prevD = -1.0;
        long long prevI = -1;
        for (int r = 0; r < k; ++r) {
            double bd = INFINITY;
            long long bi = -1;
            forKeys(g, start, cube, inner, lane, [&](unsigned p, bool have) {
                if (!have) return;
                if (!exactOnly) {
                    const float4 ka = __ldg(sorted + 2ull * p), kb = __ldg(sorted + 2ull * p + 1);
                    if (!(gridDist32<kMetric>(q32, ka, kb, so3w32, l2w32) <= limit)) return;
                }
                const unsigned ki = __ldg(perm + p);
                double key[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) key[i] = i < dim ? __ldg(keys + (unsigned long long)dim * ki + i) : 0.0;
                const double d = gridDist64<kMetric>(q, key, dim, so3w, l2w);
                ++exact;
                const bool after = d > prevD || (d == prevD && (long long)ki > prevI);
                if (after && (d < bd || (d == bd && (bi < 0 || (long long)ki < bi)))) {
                    bd = d;
                    bi = (long long)ki;
                }
            });
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double ed = __shfl_xor_sync(kFullMask, bd, o);
                const long long ei = __shfl_xor_sync(kFullMask, bi, o);
                if (ei >= 0 && (bi < 0 || ed < bd || (ed == bd && ei < bi))) {
                    bd = ed;
                    bi = ei;
                }
            }
            if (lane == 0) {
                od[r] = bi < 0 ? INFINITY : bd;
                oi[r] = bi;
            }
            if (bi < 0) {
                for (int rr = r + 1 + (int)lane; rr < k; rr += 32) {
                    od[rr] = INFINITY;
                    oi[rr] = -1;
                }
                break;
            }
            prevD = bd;
            prevI = bi;
        }
    }
    if (stats) {
        exact = __reduce_add_sync(kFullMask, exact);
        if (lane == 0) {
            atomicAdd(stats, (unsigned long long)exact);
            if (nc > (unsigned)kGridCand) atomicAdd(stats + 1, 1ull);
        }
    }
}

// two ascending (distance, index) lists per query -> one; list B's indices are offset by offB
__global__ void nnMerge2Kernel(const double *__restrict__ dA, const long long *__restrict__ iA, const double *__restrict__ dB,
                               const long long *__restrict__ iB, long long offB, unsigned long long nq, int k,
                               double *__restrict__ outD, long long *__restrict__ outI) {
    const unsigned long long qi = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const double *a = dA + qi * k, *b = dB + qi * k;
    const long long *ia = iA + qi * k, *ib = iB + qi * k;
    int pa = 0, pb = 0;
    for (int r = 0; r < k; ++r) {
        const bool ha = pa < k && ia[pa] >= 0, hb = pb < k && ib[pb] >= 0;
        bool takeA;
        if (ha && hb) takeA = a[pa] < b[pb] || (a[pa] == b[pb] && ia[pa] < ib[pb] + offB);
        else takeA = ha;
        if (!ha && !hb) {
            outD[qi * k + r] = INFINITY;
            outI[qi * k + r] = -1;
        } else if (takeA) {
            outD[qi * k + r] = a[pa];
            outI[qi * k + r] = ia[pa++];
        } else {
            outD[qi * k + r] = b[pb];
            outI[qi * k + r] = ib[pb++] + offB;
        }
    }
}

}  // namespace

cudaError_t launchNnGridBounds(const double *dKeys, size_t n, int dim, int coord0, unsigned *dOut6, cudaStream_t stream) {
    // {min x, y, z; max x, y, z} as order-preserving bit patterns
    static const unsigned init[6] = {~0u, ~0u, ~0u, 0u, 0u, 0u};
    cudaError_t e = cudaMemcpyAsync(dOut6, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    nnGridBoundsKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dKeys, (unsigned)n, dim, coord0, dOut6);
    return cudaGetLastError();
}

float nnGridOrderedToFloat(unsigned u) {
    const unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    std::memcpy(&f, &b, 4);
    return f;
}

cudaError_t launchNnGridBuild(const double *dKeys, size_t n, const NnGridGeom &g, unsigned *dCellOf, unsigned *dCounts,
                              unsigned *dStart, unsigned *dPerm, float4 *dSorted, cudaStream_t stream) {
    const size_t cells = (size_t)g.dims[0] * g.dims[1] * g.dims[2];
    cudaError_t e = cudaMemsetAsync(dCounts, 0, cells * sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    nnGridCountKernel<<<blocks, 256, 0, stream>>>(dKeys, (unsigned)n, g, dCellOf, dCounts);
    nnGridScanKernel<<<1, 1024, 0, stream>>>(dCounts, (unsigned)cells, dStart);
    e = cudaMemsetAsync(dCounts, 0, cells * sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    nnGridScatterKernel<<<blocks, 256, 0, stream>>>(dKeys, (unsigned)n, g.dim, dCellOf, dStart, dCounts, dPerm, dSorted);
    return cudaGetLastError();
}

cudaError_t launchNnGridSearch(const NnGridGeom &g, const unsigned *dStart, const unsigned *dPerm, const float4 *dSorted,
                               const double *dKeys, size_t nBuilt, double so3w, double l2w, float keysAbsMax,
                               const double *dQueries, size_t nq, int k, double *dDist, long long *dIdx,
                               unsigned long long *dStats, cudaStream_t stream) {
    if (nq == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((nq + kGridWarps - 1) / kGridWarps);
    if (g.metric == 0)
        nnGridSearchKernel<0><<<blocks, kGridWarps * 32, 0, stream>>>(g, dStart, dPerm, dSorted, dKeys, (unsigned)nBuilt, so3w, l2w,
                                                                    keysAbsMax, dQueries, nq, k, dDist, dIdx,
                                                                    (unsigned long long)k, dStats);
    else
        nnGridSearchKernel<1><<<blocks, kGridWarps * 32, 0, stream>>>(g, dStart, dPerm, dSorted, dKeys, (unsigned)nBuilt, so3w, l2w,
                                                                    keysAbsMax, dQueries, nq, k, dDist, dIdx,
                                                                    (unsigned long long)k, dStats);
    return cudaGetLastError();
}

cudaError_t launchNnMerge2(const double *dA, const long long *iA, const double *dB, const long long *iB, long long offB,
                           size_t nq, int k, double *outD, long long *outI, cudaStream_t stream) {
    if (nq == 0) return cudaSuccess;
    nnMerge2Kernel<<<(unsigned)((nq + 127) / 128), 128, 0, stream>>>(dA, iA, dB, iB, offB, nq, k, outD, outI);
    return cudaGetLastError();
}

}  // namespace mplb
