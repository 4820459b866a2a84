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
    cudaStream_t stream = nullptr;
    DevBuf queries, partDist, partIdx, dist, idx, partVal, thr;

    ~mplb_nn() {
        if (dKeys) cudaFree(dKeys);
        if (dKeys32) cudaFree(dKeys32);
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
    if (cudaStreamCreateWithFlags(&nn->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete nn;
        return fail(MPLB_ERR_CUDA, "cudaStreamCreate failed");
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
    if (nn->count + n > nn->capacity) {
        size_t cap = std::max<size_t>(1024, nn->capacity);
        while (cap < nn->count + n) cap *= 2;
        double *p = nullptr;
        cudaError_t e = cudaMalloc((void **)&p, cap * nn->dim * 8);
        if (e != cudaSuccess) return fail(MPLB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", cap * nn->dim * 8, cudaGetErrorString(e));
        if (nn->count) MPLB_CUDA(cudaMemcpy(p, nn->dKeys, nn->count * nn->dim * 8, cudaMemcpyDeviceToDevice));
        if (nn->dKeys) cudaFree(nn->dKeys);
        nn->dKeys = p;
        float *p32 = nullptr;
        e = cudaMalloc((void **)&p32, cap * nn->dim * 4);
        if (e != cudaSuccess) return fail(MPLB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", cap * nn->dim * 4, cudaGetErrorString(e));
        if (nn->count) MPLB_CUDA(cudaMemcpy(p32, nn->dKeys32, nn->count * nn->dim * 4, cudaMemcpyDeviceToDevice));
        if (nn->dKeys32) cudaFree(nn->dKeys32);
        nn->dKeys32 = p32;
        nn->capacity = cap;
    }
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

}  // extern "C"
