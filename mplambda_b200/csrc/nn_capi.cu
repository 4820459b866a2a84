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

    ~mplb_nn() {
        if (dKeys) cudaFree(dKeys);
        if (dKeys32) cudaFree(dKeys32);
        if (hKeyStats) cudaFreeHost(hKeyStats);
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
    const int slices = nnSlices(nq, nn->count, nn->numSms);
    int rc;
    if (!deviceQueries && (rc = nnSyncMirror(nn))) return rc;
    MPLB_CUDA(cudaStreamWaitEvent(st, nn->scratchDone, 0));
    if (slices > 1 && ((rc = nn->partDist.reserve(nq * slices * k * 8)) || (rc = nn->partIdx.reserve(nq * slices * k * 8))))
        return rc;
    // k == 1: [nq][slices] bounds; k > 1: [nq][k] sample distances + [nq][8] counts + [nq] chosen levels
    if ((rc = nn->partVal.reserve(std::max<size_t>(nq * slices * k * 4, nq * ((size_t)k + 17) * 4))) || (rc = nn->thr.reserve(nq * 4)))
        return rc;
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
    MPLB_CUDA(launchNnSearch(nn->dKeys, nn->dKeys32, (std::isfinite(nn->keysAbsMax) ? std::nextafterf((float)nn->keysAbsMax, INFINITY) : INFINITY), nn->count, nn->dim,
                             nn->metric, nn->so3w, nn->l2w, dQ, nq, k, slices, (float *)nn->partVal.ptr,
                             (float *)nn->thr.ptr, (double *)nn->partDist.ptr, (long long *)nn->partIdx.ptr, dDist, dIdx,
                             nullptr, st));
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
    *out = nn;
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
    const int slices = nnSlices(nq, nn->count, nn->numSms);
    int rc;
    if (slices > 1 && ((rc = nn->partDist.reserve(nq * slices * k * 8)) || (rc = nn->partIdx.reserve(nq * slices * k * 8))))
        return rc;
    if ((rc = nn->partVal.reserve(std::max<size_t>(nq * slices * k * 4, nq * ((size_t)k + 17) * 4))) || (rc = nn->thr.reserve(nq * 4)))
        return rc;
    MPLB_CUDA(cudaStreamWaitEvent(st, nn->scratchDone, 0));
    MPLB_CUDA(launchNnSearch(nn->dKeys, nn->dKeys32, (std::isfinite(nn->keysAbsMax) ? std::nextafterf((float)nn->keysAbsMax, INFINITY) : INFINITY), nn->count, nn->dim,
                             nn->metric, nn->so3w, nn->l2w, dQueries, nq, k, slices, (float *)nn->partVal.ptr,
                             (float *)nn->thr.ptr, (double *)nn->partDist.ptr, (long long *)nn->partIdx.ptr, dDist, dIdx,
                             nullptr, st));
    MPLB_CUDA(cudaEventRecord(nn->scratchDone, st));
    return MPLB_OK;
}
int nnInfo(const mplb_nn *nn, int *device, int *metric, int *dim, double *so3w, double *l2w) {
    *device = nn->device; *metric = nn->metric; *dim = nn->dim; *so3w = nn->so3w; *l2w = nn->l2w;
    return MPLB_OK;
}
}  // namespace mplb
