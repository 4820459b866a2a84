// C-ABI of the nearest-neighbour structure (include/mplb.h, mplb_nn_*): replaces nigh::Nigh as the
// planners use it (reference include/mpl/prrt.hpp:42,61,67,115; include/mpl/pcforest.hpp:51,81,86,107,193).
#include "../../include/mplb.h"
#include "nn_device.cuh"
#include "scene_impl.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

using namespace mplb;

struct mplb_nn {
    int device = 0, numSms = 0, metric = 0, dim = 7;
    double so3w = 50, l2w = 1;
    std::mutex mu;                 // calls on one structure are serialised
    size_t count = 0, capacity = 0;
    double *dKeys = nullptr;
    float *dKeys32 = nullptr;      // FP32 copy for the candidate pass
    double keysAbsMax = 0;         // largest |translation coordinate| (SE3) / |coordinate| (L1) inserted
    std::vector<double> hKeys;     // host mirror: reported distances are recomputed with host libm
    size_t hCount = 0;             // keys mirrored so far (keys inserted from device memory are fetched on demand)
    cudaStream_t stream = nullptr;
    // the scan scratch below is shared by every query on this structure, whatever stream it runs on: each search waits
    // for the previous one's event before it touches the scratch and records its own behind its last kernel
    cudaEvent_t scratchDone = nullptr;
    float *hKeyStats = nullptr;    // pinned: {largest |coordinate|, non-unit quaternion seen} of a device-side insert
    DevBuf queries, partDist, partIdx, dist, idx, partVal, thr, keyStats;
    // Spatial index over the translations of SE(3) keys (nn_grid.cu): keys [0, gridBuilt) are indexed, the ones inserted
    // since are searched by the brute-force kernels and the two lists merged; rebuilt when that tail outgrows a quarter
    // of the index.  gridOn: mplb_nn_set_index (default: on from kGridMin keys); gridTried: key count of the last
    // attempt that found nothing worth indexing (all translations in a handful of cells)
    bool gridOn = true;
    size_t gridBuilt = 0, gridTried = 0, gridCells = 0;
    NnGridGeom grid{};
    unsigned *hGridBounds = nullptr;   // pinned
    DevBuf gridBounds, gridStart, gridPerm, gridSorted, gridCellOf, gridCounts, gridDist, gridIdx, tailDist, tailIdx, gridStats;

    ~mplb_nn() {
        if (dKeys) cudaFree(dKeys);
        if (dKeys32) cudaFree(dKeys32);
        if (hKeyStats) cudaFreeHost(hKeyStats);
        if (hGridBounds) cudaFreeHost(hGridBounds);
        if (scratchDone) cudaEventDestroy(scratchDone);
        if (stream) cudaStreamDestroy(stream);
    }
    // nigh SE3Space / L1Space distance, operation order of oracle/mplb_oracle.c
    double distance(const double *a, const double *b) const {
        if (metric == 0) {
            double d = std::fabs(((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) + a[3] * b[3]);
            double so3 = d < 1.0 ? std::acos(d) : 0.0;
            double dx = a[4] - b[4], dy = a[5] - b[5], dz = a[6] - b[6];
            return so3w * so3 + l2w * std::sqrt((dx * dx + dy * dy) + dz * dz);
        }
        double s = 0;
        for (int i = 0; i < dim; ++i) s += std::fabs(a[i] - b[i]);
        return s;
    }
};

// room for `total` keys; existing keys move with the arrays (device pointers handed out before are stale afterwards)
static int nnEnsureCapacity(mplb_nn *nn, size_t total) {
    if (total <= nn->capacity) return MPLB_OK;
    size_t cap = std::max<size_t>(1024, nn->capacity);
    while (cap < total) cap *= 2;
    MPLB_CUDA(cudaDeviceSynchronize());   // (rare: the arrays double) nothing may still be reading or filling the old ones
    double *p = nullptr;
    float *p32 = nullptr;
    cudaError_t e = cudaMalloc((void **)&p, cap * nn->dim * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p32, cap * nn->dim * 4);
    if (e == cudaSuccess && nn->count) e = cudaMemcpy(p, nn->dKeys, nn->count * nn->dim * 8, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess && nn->count) e = cudaMemcpy(p32, nn->dKeys32, nn->count * nn->dim * 4, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        if (p32) cudaFree(p32);
        return fail(MPLB_ERR_NOMEM, "growing the key arrays to %zu keys failed: %s", cap, cudaGetErrorString(e));
    }
    if (nn->dKeys) cudaFree(nn->dKeys);
    if (nn->dKeys32) cudaFree(nn->dKeys32);
    nn->dKeys = p;
    nn->dKeys32 = p32;
    nn->capacity = cap;
    return MPLB_OK;
}

// host copies of the keys that were inserted from device memory (the host-facing queries report host-libm distances)
static int nnSyncMirror(mplb_nn *nn) {
    if (nn->hCount == nn->count) return MPLB_OK;
    nn->hKeys.resize(nn->count * nn->dim);
    MPLB_CUDA(cudaMemcpy(nn->hKeys.data() + nn->hCount * nn->dim, nn->dKeys + nn->hCount * nn->dim,
                         (nn->count - nn->hCount) * nn->dim * 8, cudaMemcpyDeviceToHost));
    nn->hCount = nn->count;
    return MPLB_OK;
}

static void nnNoteKeyStats(mplb_nn *nn, double absMax, bool nonUnit) {
    if (nonUnit || !std::isfinite(absMax)) nn->keysAbsMax = INFINITY;
    else if (std::isfinite(nn->keysAbsMax)) nn->keysAbsMax = std::max(nn->keysAbsMax, absMax);
}

// ---- spatial index ---------------------------------------------------------------------------------------------
static size_t nnGridMin() {
    static const size_t v = [] {
        const char *e = std::getenv("MPLB_NN_GRID_MIN");
        return e ? (size_t)std::max(1L, std::atol(e)) : (size_t)256;
    }();
    return v;
}

// (re)builds the index over all keys there are, on `st`; leaves gridBuilt == 0 when the keys do not spread over enough
// cells to be worth it
static int nnGridRebuild(mplb_nn *nn, cudaStream_t st) {
    const size_t n = nn->count;
    nn->gridTried = n;
    int rc;
    if (!nn->hGridBounds && cudaMallocHost((void **)&nn->hGridBounds, 6 * sizeof(unsigned)) != cudaSuccess)
        return fail(MPLB_ERR_NOMEM, "pinned scratch for the index bounds");
    if ((rc = nn->gridBounds.reserve(6 * sizeof(unsigned))) || (rc = nn->gridStats.reserve(16))) return rc;
    const int coord0 = nn->metric == 0 ? 4 : 0;
    MPLB_CUDA(launchNnGridBounds(nn->dKeys, n, nn->dim, coord0, (unsigned *)nn->gridBounds.ptr, st));
    MPLB_CUDA(cudaMemcpyAsync(nn->hGridBounds, nn->gridBounds.ptr, 6 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    MPLB_CUDA(cudaStreamSynchronize(st));
    double lo[3], ext[3], maxExt = 0;
    for (int a = 0; a < 3; ++a) {
        lo[a] = nnGridOrderedToFloat(nn->hGridBounds[a]);
        ext[a] = (double)nnGridOrderedToFloat(nn->hGridBounds[3 + a]) - lo[a];
        if (!(ext[a] >= 0) || !std::isfinite(ext[a])) return MPLB_OK;   // (non-finite keys never get here; be safe)
        maxExt = std::max(maxExt, ext[a]);
    }
    int nd = 0;
    double vol = 1;
    for (int a = 0; a < 3; ++a)
        if (ext[a] > 1e-9 * maxExt && ext[a] > 0) {
            vol *= ext[a];
            ++nd;
        }
    nn->gridBuilt = 0;
    if (nd == 0) return MPLB_OK;
    // ~8 keys per cell on average; at most 1024 cells per axis and 2^22 in all
    double cells = std::min((double)n / 8.0, 4194304.0);
    double cell = std::pow(vol / cells, 1.0 / nd);
    NnGridGeom g{};
    for (int tries = 0; tries < 8; ++tries) {
        size_t total = 1;
        for (int a = 0; a < 3; ++a) {
            g.dims[a] = (int)std::min(1024.0, std::max(1.0, std::ceil(ext[a] / cell * 1.0000001 + 1e-9)));
            total *= (size_t)g.dims[a];
        }
        if (total <= 4194304) break;
        cell *= 1.3;
    }
    const size_t total = (size_t)g.dims[0] * g.dims[1] * g.dims[2];
    if (total < 64 || total > 4194304) return MPLB_OK;
    for (int a = 0; a < 3; ++a) {
        // a cell wide enough for the axis even when the 1024-cell cap is what set dims
        cell = std::max(cell, ext[a] / g.dims[a] * 1.0000001);
    }
    for (int a = 0; a < 3; ++a) g.origin[a] = lo[a] - 1e-6 * cell;
    g.cell = cell;
    g.invCell = 1.0 / cell;
    g.metric = nn->metric;
    g.dim = nn->dim;
    g.coord0 = coord0;
    if ((rc = nn->gridStart.reserve((total + 1) * 4)) || (rc = nn->gridCounts.reserve(total * 4)) ||
        (rc = nn->gridPerm.reserve(n * 4)) || (rc = nn->gridCellOf.reserve(n * 4)) || (rc = nn->gridSorted.reserve(n * 32)))
        return rc;
    MPLB_CUDA(launchNnGridBuild(nn->dKeys, n, g, (unsigned *)nn->gridCellOf.ptr, (unsigned *)nn->gridCounts.ptr,
                                (unsigned *)nn->gridStart.ptr, (unsigned *)nn->gridPerm.ptr, (float4 *)nn->gridSorted.ptr, st));
    nn->grid = g;
    nn->gridBuilt = n;
    nn->gridCells = total;
    return MPLB_OK;
}

// The search itself on `st`, every pointer device memory: through the index (plus the brute-force kernels over the keys
// inserted since it was built) or by brute force alone.  The caller holds nn->mu and has ordered `st` behind the
// previous search (the scratch is shared).
static int nnLaunchSearch(mplb_nn *nn, const double *dQ, size_t nq, int k, double *dDist, long long *dIdx, cudaStream_t st) {
    int rc;
    const float absMax = std::isfinite(nn->keysAbsMax) ? std::nextafterf((float)nn->keysAbsMax, INFINITY) : INFINITY;
    const bool gridOk = nn->gridOn && (nn->metric == 0 || nn->dim >= 3) && std::isfinite(nn->keysAbsMax) && nn->count >= nnGridMin() &&
                        nn->count < ((size_t)1 << 31);
    if (!gridOk) nn->gridBuilt = 0;
    if (gridOk) {
        const size_t since = nn->count - std::max(nn->gridBuilt, nn->gridTried);
        // Rebuilding is three small kernels, a scan and one 24-byte read-back (~50 us for a few thousand keys, ~0.15 ms
        // for 65,536); searching the keys inserted since by brute force costs 0.1-0.9 ms whatever their number (that
        // pipeline's fixed cost).  So a structure of up to 65,536 keys is re-indexed whenever keys have come in, a larger
        // one when they outgrow a quarter of the index.
        const size_t tail = nn->count - nn->gridBuilt;
        if ((nn->gridBuilt == 0 && (nn->gridTried == 0 || since > nn->gridTried / 4)) ||
            (nn->gridBuilt != 0 && tail > (nn->count <= 65536 ? (size_t)0 : std::max<size_t>(4096, nn->gridBuilt / 4))))
            if ((rc = nnGridRebuild(nn, st))) return rc;
    }
    auto brute = [&](size_t first, double *oDist, long long *oIdx) -> int {
        const size_t n = nn->count - first;
        const int slices = nnSlices(nq, n, nn->numSms);
        int r;
        if (slices > 1 && ((r = nn->partDist.reserve(nq * slices * k * 8)) || (r = nn->partIdx.reserve(nq * slices * k * 8))))
            return r;
        // k == 1: [nq][slices] bounds; k > 1: [nq][k] sample distances + [nq][8] counts + [nq] chosen levels
        if ((r = nn->partVal.reserve(std::max<size_t>(nq * slices * k * 4, nq * ((size_t)k + 17) * 4))) || (r = nn->thr.reserve(nq * 4)))
            return r;
        MPLB_CUDA(launchNnSearch(nn->dKeys + first * nn->dim, nn->dKeys32 + first * nn->dim, absMax, n, nn->dim, nn->metric,
                                 nn->so3w, nn->l2w, dQ, nq, k, slices, (float *)nn->partVal.ptr, (float *)nn->thr.ptr,
                                 (double *)nn->partDist.ptr, (long long *)nn->partIdx.ptr, oDist, oIdx, nullptr, st));
        return MPLB_OK;
    };
    if (!gridOk || nn->gridBuilt == 0) return brute(0, dDist, dIdx);
    const size_t tail = nn->count - nn->gridBuilt;
    double *gd = dDist;
    long long *gi = dIdx;
    if (tail) {
        if ((rc = nn->gridDist.reserve(nq * k * 8)) || (rc = nn->gridIdx.reserve(nq * k * 8)) ||
            (rc = nn->tailDist.reserve(nq * k * 8)) || (rc = nn->tailIdx.reserve(nq * k * 8)))
            return rc;
        gd = (double *)nn->gridDist.ptr;
        gi = (long long *)nn->gridIdx.ptr;
    }
    MPLB_CUDA(launchNnGridSearch(nn->grid, (const unsigned *)nn->gridStart.ptr, (const unsigned *)nn->gridPerm.ptr,
                                 (const float4 *)nn->gridSorted.ptr, nn->dKeys, nn->gridBuilt, nn->so3w, nn->l2w, absMax, dQ, nq,
                                 k, gd, gi, (unsigned long long *)nn->gridStats.ptr, st));
    if (!tail) return MPLB_OK;
    if ((rc = brute(nn->gridBuilt, (double *)nn->tailDist.ptr, (long long *)nn->tailIdx.ptr))) return rc;
    MPLB_CUDA(launchNnMerge2(gd, gi, (const double *)nn->tailDist.ptr, (const long long *)nn->tailIdx.ptr,
                             (long long)nn->gridBuilt, nq, k, dDist, dIdx, st));
    return MPLB_OK;
}

static int nnSearch(mplb_nn *nn, const double *queries, bool deviceQueries, size_t nq, int k, int64_t *idx,
                    double *dist, bool deviceOut, cudaStream_t userStream) {
    if (k < 1 || k > 64) return fail(MPLB_ERR_INVALID, "k must be in [1, 64], got %d", k);
    std::lock_guard<std::mutex> g(nn->mu);
    MPLB_CUDA(cudaSetDevice(nn->device));
    cudaStream_t st = deviceQueries ? userStream : nn->stream;
    if (nn->count == 0) {  // nearest() on an empty structure: no result (nigh returns an empty optional)
        if (deviceOut) {
            std::vector<int64_t> hi(nq * k, -1);
            std::vector<double> hd(nq * k, INFINITY);
            MPLB_CUDA(cudaMemcpyAsync(idx, hi.data(), hi.size() * 8, cudaMemcpyHostToDevice, st));
            MPLB_CUDA(cudaMemcpyAsync(dist, hd.data(), hd.size() * 8, cudaMemcpyHostToDevice, st));
            MPLB_CUDA(cudaStreamSynchronize(st));
        } else {
            for (size_t i = 0; i < nq * k; ++i) {
                idx[i] = -1;
                dist[i] = INFINITY;
            }
        }
        return MPLB_OK;
    }
    int rc;
    if (!deviceQueries && (rc = nnSyncMirror(nn))) return rc;
    MPLB_CUDA(cudaStreamWaitEvent(st, nn->scratchDone, 0));
    const double *dQ = queries;
    double *dDist = dist;
    long long *dIdx = (long long *)idx;
    if (!deviceQueries) {
        if ((rc = nn->queries.reserve(nq * nn->dim * 8)) || (rc = nn->dist.reserve(nq * k * 8)) ||
            (rc = nn->idx.reserve(nq * k * 8)))
            return rc;
        MPLB_CUDA(cudaMemcpyAsync(nn->queries.ptr, queries, nq * nn->dim * 8, cudaMemcpyHostToDevice, st));
        dQ = (const double *)nn->queries.ptr;
        dDist = (double *)nn->dist.ptr;
        dIdx = (long long *)nn->idx.ptr;
    }
    if ((rc = nnLaunchSearch(nn, dQ, nq, k, dDist, dIdx, st))) return rc;
    MPLB_CUDA(cudaEventRecord(nn->scratchDone, st));
    if (deviceQueries) return MPLB_OK;
    MPLB_CUDA(cudaMemcpyAsync(idx, dIdx, nq * k * 8, cudaMemcpyDeviceToHost, st));
    MPLB_CUDA(cudaStreamSynchronize(st));
    // distances as the CPU path would report them (host libm, same operation order); nq x k acos/sqrt calls are
    // milliseconds on one core, so large batches use all of them
#pragma omp parallel for schedule(static) if (nq * (size_t)k >= 16384)
    for (long long i = 0; i < (long long)nq; ++i)
        for (int r = 0; r < k; ++r) {
            const int64_t j = idx[i * k + r];
            dist[i * k + r] = j < 0 ? INFINITY : nn->distance(nn->hKeys.data() + (size_t)j * nn->dim, queries + i * nn->dim);
        }
    return MPLB_OK;
}

extern "C" {

int mplb_nn_create(int metric, int dim, double so3_weight, double l2_weight, int device, mplb_nn **out) {
    if (!out) return fail(MPLB_ERR_INVALID, "out is null");
    *out = nullptr;
    if (metric != 0 && metric != 1) return fail(MPLB_ERR_INVALID, "metric must be 0 (SE3) or 1 (L1)");
    if (metric == 0) dim = 7;
    if (dim < 1 || dim > 8) return fail(MPLB_ERR_INVALID, "dim must be in [1, 8]");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(MPLB_ERR_CUDA, "no CUDA device available (%s); libmplb has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(MPLB_ERR_INVALID, "device %d out of range [0,%d)", device, count);
    MPLB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MPLB_CUDA(cudaGetDeviceProperties(&prop, device));
    auto *nn = new mplb_nn();
    nn->device = device;
    nn->numSms = prop.multiProcessorCount;
    nn->metric = metric;
    nn->dim = dim;
    nn->so3w = so3_weight;
    nn->l2w = l2_weight;
    if (cudaStreamCreateWithFlags(&nn->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&nn->scratchDone, cudaEventDisableTiming) != cudaSuccess ||
        cudaMallocHost((void **)&nn->hKeyStats, 2 * sizeof(float)) != cudaSuccess || nn->keyStats.reserve(2 * sizeof(float))) {
        delete nn;
        return fail(MPLB_ERR_CUDA, "could not create the structure's stream / event / pinned scratch");
    }
    if (nn->gridStats.reserve(16) || cudaMemset(nn->gridStats.ptr, 0, 16) != cudaSuccess) {
        delete nn;
        return fail(MPLB_ERR_CUDA, "could not create the structure's index counters");
    }
    *out = nn;
    return MPLB_OK;
}

int mplb_nn_set_index(mplb_nn *nn, int on) {
    if (!nn) return fail(MPLB_ERR_INVALID, "nn is null");
    std::lock_guard<std::mutex> g(nn->mu);
    nn->gridOn = on != 0;
    if (!on) nn->gridBuilt = nn->gridTried = 0;
    return MPLB_OK;
}

int mplb_nn_index_info(mplb_nn *nn, size_t *indexed, size_t *cells, uint64_t *exact_evaluations, uint64_t *slow_queries) {
    if (!nn) return fail(MPLB_ERR_INVALID, "nn is null");
    std::lock_guard<std::mutex> g(nn->mu);
    MPLB_CUDA(cudaSetDevice(nn->device));
    unsigned long long st[2] = {0, 0};
    MPLB_CUDA(cudaDeviceSynchronize());
    MPLB_CUDA(cudaMemcpy(st, nn->gridStats.ptr, sizeof(st), cudaMemcpyDeviceToHost));
    if (indexed) *indexed = nn->gridBuilt;
    if (cells) *cells = nn->gridBuilt ? nn->gridCells : 0;
    if (exact_evaluations) *exact_evaluations = st[0];
    if (slow_queries) *slow_queries = st[1];
    return MPLB_OK;
}

void mplb_nn_destroy(mplb_nn *nn) {
    if (!nn) return;
    cudaSetDevice(nn->device);
    cudaDeviceSynchronize();
    delete nn;
}

int mplb_nn_insert(mplb_nn *nn, const double *keys, size_t n) {
    if (!nn) return fail(MPLB_ERR_INVALID, "nn is null");
    if (n == 0) return MPLB_OK;
    if (!keys) return fail(MPLB_ERR_INVALID, "null keys");
    std::lock_guard<std::mutex> g(nn->mu);
    MPLB_CUDA(cudaSetDevice(nn->device));
    int rc;
    if ((rc = nnEnsureCapacity(nn, nn->count + n)) || (rc = nnSyncMirror(nn))) return rc;
    for (size_t i = 0; i < n; ++i)
        for (int d = (nn->metric == 0 ? 4 : 0); d < nn->dim; ++d) {
            const double v = std::fabs(keys[i * nn->dim + d]);
            if (!std::isfinite(v)) return fail(MPLB_ERR_INVALID, "non-finite key %zu", i);
            if (std::isfinite(nn->keysAbsMax)) nn->keysAbsMax = std::max(nn->keysAbsMax, v);
        }
    if (nn->metric == 0)
        for (size_t i = 0; i < n; ++i) {   // non-unit quaternion keys: the FP32 candidate bound does not hold -> all exact
            const double *q = keys + i * 7;
            const double qn = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
            if (!(std::fabs(qn - 1.0) < 1e-3)) nn->keysAbsMax = INFINITY;
        }
    // on the structure's own stream (queries are ordered behind it); no device-wide synchronisation: other threads'
    // scenes keep running
    MPLB_CUDA(cudaMemcpyAsync(nn->dKeys + nn->count * nn->dim, keys, n * nn->dim * 8, cudaMemcpyHostToDevice, nn->stream));
    MPLB_CUDA(launchNnToFloat(nn->dKeys + nn->count * nn->dim, nn->dKeys32 + nn->count * nn->dim, n * nn->dim, nn->stream));
    MPLB_CUDA(cudaStreamSynchronize(nn->stream));
    nn->hKeys.insert(nn->hKeys.end(), keys, keys + n * nn->dim);
    nn->count += n;
    nn->hCount = nn->count;
    return MPLB_OK;
}

size_t mplb_nn_size(const mplb_nn *nn) { return nn ? nn->count : 0; }

int mplb_nn_nearest(mplb_nn *nn, const double *queries, size_t nq, int64_t *idx, double *dist) {
    if (!nn) return fail(MPLB_ERR_INVALID, "nn is null");
    if (nq == 0) return MPLB_OK;
    if (!queries || !idx || !dist) return fail(MPLB_ERR_INVALID, "null buffer");
    return nnSearch(nn, queries, false, nq, 1, idx, dist, false, nullptr);
}

int mplb_nn_knearest(mplb_nn *nn, const double *queries, size_t nq, int k, int64_t *idx, double *dist) {
    if (!nn) return fail(MPLB_ERR_INVALID, "nn is null");
    if (nq == 0) return MPLB_OK;
    if (!queries || !idx || !dist) return fail(MPLB_ERR_INVALID, "null buffer");
    return nnSearch(nn, queries, false, nq, k, idx, dist, false, nullptr);
}

int mplb_nn_nearest_device(mplb_nn *nn, const double *d_queries, size_t nq, int64_t *d_idx, double *d_dist, void *stream) {
    if (!nn) return fail(MPLB_ERR_INVALID, "nn is null");
    if (nq == 0) return MPLB_OK;
    if (!d_queries || !d_idx || !d_dist) return fail(MPLB_ERR_INVALID, "null buffer");
    return nnSearch(nn, d_queries, true, nq, 1, d_idx, d_dist, true, (cudaStream_t)stream);
}

int mplb_nn_knearest_device(mplb_nn *nn, const double *d_queries, size_t nq, int k, int64_t *d_idx, double *d_dist,
                            void *stream) {
    if (!nn) return fail(MPLB_ERR_INVALID, "nn is null");
    if (nq == 0) return MPLB_OK;
    if (!d_queries || !d_idx || !d_dist) return fail(MPLB_ERR_INVALID, "null buffer");
    return nnSearch(nn, d_queries, true, nq, k, d_idx, d_dist, true, (cudaStream_t)stream);
}

int mplb_nn_reserve(mplb_nn *nn, size_t capacity) {
    if (!nn) return fail(MPLB_ERR_INVALID, "nn is null");
    std::lock_guard<std::mutex> g(nn->mu);
    MPLB_CUDA(cudaSetDevice(nn->device));
    return nnEnsureCapacity(nn, capacity);
}

size_t mplb_nn_capacity(const mplb_nn *nn) { return nn ? nn->capacity : 0; }

const double *mplb_nn_keys_device(const mplb_nn *nn) { return nn ? nn->dKeys : nullptr; }

int mplb_nn_insert_device(mplb_nn *nn, const double *d_keys, size_t n, void *stream) {
    if (!nn) return fail(MPLB_ERR_INVALID, "nn is null");
    if (n == 0) return MPLB_OK;
    if (!d_keys) return fail(MPLB_ERR_INVALID, "null keys");
    std::lock_guard<std::mutex> g(nn->mu);
    MPLB_CUDA(cudaSetDevice(nn->device));
    int rc;
    if ((rc = nnEnsureCapacity(nn, nn->count + n))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    double *dst = nn->dKeys + nn->count * nn->dim;
    MPLB_CUDA(cudaMemcpyAsync(dst, d_keys, n * nn->dim * 8, cudaMemcpyDeviceToDevice, st));
    MPLB_CUDA(launchNnToFloat(dst, nn->dKeys32 + nn->count * nn->dim, n * nn->dim, st));
    MPLB_CUDA(launchNnKeyStats(dst, n, nn->dim, nn->metric, (float *)nn->keyStats.ptr, st));
    MPLB_CUDA(cudaMemcpyAsync(nn->hKeyStats, nn->keyStats.ptr, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    MPLB_CUDA(cudaStreamSynchronize(st));
    if (!(nn->hKeyStats[0] < INFINITY)) return fail(MPLB_ERR_INVALID, "non-finite key in a device-side insert");
    nnNoteKeyStats(nn, nn->hKeyStats[0], nn->hKeyStats[1] != 0.f);
    nn->count += n;
    return MPLB_OK;
}

}  // extern "C"

// ---- used by the device-resident planner loop (rrt_capi.cu): keys appended in place by a kernel
namespace mplb {
int nnLock(mplb_nn *nn, size_t room, double **keys, float **keys32, size_t *count) {
    nn->mu.lock();
    int rc = nnEnsureCapacity(nn, nn->count + room);
    if (rc != MPLB_OK) {
        nn->mu.unlock();
        return rc;
    }
    *keys = nn->dKeys;
    *keys32 = nn->dKeys32;
    *count = nn->count;
    return MPLB_OK;
}
void nnUnlock(mplb_nn *nn, size_t added, double absMaxBound) {
    if (added) nnNoteKeyStats(nn, absMaxBound, false);
    nn->count += added;
    nn->mu.unlock();
}
int nnSearchDeviceLocked(mplb_nn *nn, const double *dQueries, size_t nq, int k, long long *dIdx, double *dDist,
                         cudaStream_t st) {
    if (nn->count == 0) return fail(MPLB_ERR_INVALID, "nearest-neighbour structure is empty");
    int rc;
    MPLB_CUDA(cudaStreamWaitEvent(st, nn->scratchDone, 0));
    if ((rc = nnLaunchSearch(nn, dQueries, nq, k, dDist, dIdx, st))) return rc;
    MPLB_CUDA(cudaEventRecord(nn->scratchDone, st));
    return MPLB_OK;
}
int nnInfo(const mplb_nn *nn, int *device, int *metric, int *dim, double *so3w, double *l2w) {
    *device = nn->device; *metric = nn->metric; *dim = nn->dim; *so3w = nn->so3w; *l2w = nn->l2w;
    return MPLB_OK;
}
}  // namespace mplb
