// C-ABI of the device-resident planner iteration (include/mplb.h: mplb_sample_states_device,
// mplb_check_edges_indexed*, mplb_rrt_*): sample -> nearest -> motion check -> insert without a state crossing PCIe.
// Reference loop: Planner<Scenario, PRRT>::Thread::addSample (include/mpl/prrt.hpp:280-299); the index-named motion
// check is what PCForest's addNodeNear issues for every neighbour (include/mpl/pcforest.hpp:507-567).
#include "../../include/mplb.h"
#include "planner_device.cuh"
#include "scene_impl.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

struct mplb_nn;

namespace mplb {
// nn_capi.cu
int nnLock(mplb_nn *nn, size_t room, double **keys, float **keys32, size_t *count);
void nnUnlock(mplb_nn *nn, size_t added, double absMaxBound);
int nnSearchDeviceLocked(mplb_nn *nn, const double *dQueries, size_t nq, int k, long long *dIdx, double *dDist, cudaStream_t st);
int nnInfo(const mplb_nn *nn, int *device, int *metric, int *dim, double *so3w, double *l2w);
// capi.cu
int se3CheckEdgesHost(Se3Scene *sc, const double *from, const double *to, size_t n, uint8_t *valid, uint64_t *statesChecked);

namespace {

Se3Scene *se3Of(mplb_scene *scene) { return dynamic_cast<Se3Scene *>(reinterpret_cast<SceneBase *>(scene)); }

// Index-named motions on the context's stream: plan + check, verdicts left in c->out (device), counters in c->ctr.
int enqueueIndexed(Se3Scene *sc, CallCtx *c, EdgeGather g, size_t n) {
    int rc;
    const size_t stackBytes = edgeStackBytes(sc->numSms);
    if ((rc = c->dead.reserve(n * 4)) || (rc = c->out.reserve(n)) || (rc = c->aux0.reserve(stackBytes)) ||
        (rc = c->aux1.reserve(edgeGatherBytes(n))))
        return rc;
    static const unsigned force = [] { const char *e = std::getenv("MPLB_DEBUG_AMBIGUOUS_EVERY"); return e ? (unsigned)std::atoi(e) : 0u; }();
    g.forceAmbiguousEvery = force;
    MPLB_CUDA(cudaMemsetAsync(c->ctr.ptr, 0, sizeof(DevCounters), c->stream));
    MPLB_CUDA(launchCheckEdgesIndexed(sc->dev, g, n, sc->so3w, sc->l2w, sc->invStep, (unsigned *)c->dead.ptr, (uint8_t *)c->out.ptr,
                                      (DevCounters *)c->ctr.ptr, (unsigned long long *)c->work.ptr, c->aux0.ptr, stackBytes,
                                      c->aux1.ptr, sc->numSms, c->stream));
    sc->hLaunches += 6;   // gather + plan, three passes, per-edge verdicts, (memsets not counted)
    return MPLB_OK;
}

// After the stream has drained and c->hCtr holds the call's counters: edges whose device-planned step count may differ
// from the host's (CUDA's acos against glibc's, se3_kernels.cu: edgeGatherPlanKernel) are planned again with host libm
// and, where the counts really differ, checked again through the host path; their verdicts replace the ones in c->out.
// -> number of verdicts replaced (< 0: error code)
int resolveIndexed(Se3Scene *sc, CallCtx *c, size_t n) {
    const DevCounters &k = *c->hCtr;
    if (k.badEdges) return fail(MPLB_ERR_INVALID, "%u edges need >= 2^31 interpolation steps; non-finite state?", k.badEdges);
    if (!k.ambiguous) return 0;
    sc->hAmbiguous += k.ambiguous;
    std::vector<unsigned> edges;
    if (k.ambiguous <= kAmbiguousListed) edges.assign(k.ambiguousEdge, k.ambiguousEdge + k.ambiguous);
    else {
        edges.resize(n);
        for (size_t e = 0; e < n; ++e) edges[e] = (unsigned)e;
    }
    const double *dFrom = edgeGatherFrom(c->aux1.ptr, n), *dTo = edgeGatherTo(c->aux1.ptr, n);
    const unsigned *dSteps = edgeGatherSteps(c->aux1.ptr, n);
    std::vector<double> from(7 * edges.size()), to(7 * edges.size());
    std::vector<unsigned> steps(edges.size());
    if (edges.size() == n) {
        MPLB_CUDA(cudaMemcpy(from.data(), dFrom, n * 56, cudaMemcpyDeviceToHost));
        MPLB_CUDA(cudaMemcpy(to.data(), dTo, n * 56, cudaMemcpyDeviceToHost));
        MPLB_CUDA(cudaMemcpy(steps.data(), dSteps, n * 4, cudaMemcpyDeviceToHost));
    } else {
        for (size_t i = 0; i < edges.size(); ++i) {
            MPLB_CUDA(cudaMemcpy(&from[7 * i], dFrom + 7 * (size_t)edges[i], 56, cudaMemcpyDeviceToHost));
            MPLB_CUDA(cudaMemcpy(&to[7 * i], dTo + 7 * (size_t)edges[i], 56, cudaMemcpyDeviceToHost));
            MPLB_CUDA(cudaMemcpy(&steps[i], dSteps + edges[i], 4, cudaMemcpyDeviceToHost));
        }
    }
    static const bool replanAll = std::getenv("MPLB_DEBUG_REPLAN_ALL") != nullptr;   // test hook: exercise the re-check
    std::vector<size_t> redo;
    for (size_t i = 0; i < edges.size(); ++i)
        if (replanAll || sc->edgeSteps(&from[7 * i], &to[7 * i]) != steps[i]) redo.push_back(i);
    if (redo.empty()) return 0;
    std::vector<double> rf(7 * redo.size()), rt(7 * redo.size());
    std::vector<uint8_t> rv(redo.size());
    for (size_t j = 0; j < redo.size(); ++j) {
        std::memcpy(&rf[7 * j], &from[7 * redo[j]], 56);
        std::memcpy(&rt[7 * j], &to[7 * redo[j]], 56);
    }
    int rc = se3CheckEdgesHost(sc, rf.data(), rt.data(), redo.size(), rv.data(), nullptr);
    if (rc != MPLB_OK) return rc;
    for (size_t j = 0; j < redo.size(); ++j)
        MPLB_CUDA(cudaMemcpy((uint8_t *)c->out.ptr + edges[redo[j]], &rv[j], 1, cudaMemcpyHostToDevice));
    return (int)redo.size();
}

int finishIndexed(Se3Scene *sc, CallCtx *c, size_t n, uint64_t *statesChecked) {
    MPLB_CUDA(cudaMemcpyAsync(c->hCtr, c->ctr.ptr, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
    MPLB_CUDA(cudaStreamSynchronize(c->stream));
    const DevCounters &k = *c->hCtr;
    if (k.overflow) return fail(MPLB_ERR_CUDA, "device traversal bookkeeping error (pair stack / slot accounting)");
    sc->hBv += k.bvTests;
    sc->hTri += k.triTests;
    sc->hExact += k.exactTests;
    sc->hChecks += k.states;
    if (statesChecked) *statesChecked = k.states;
    const int r = resolveIndexed(sc, c, n);
    return r < 0 ? r : MPLB_OK;
}

}  // namespace
}  // namespace mplb

using namespace mplb;

// ---- device-resident RRT (prrt.hpp:280-299 batched; the tree lives in the nearest-neighbour structure's key array)
struct mplb_rrt {
    Se3Scene *scene = nullptr;
    mplb_nn *nn = nullptr;
    CallCtx *ctx = nullptr;
    size_t batch = 0;
    SampleSe3 sampler{};
    double goalRadius = 0.1, absMaxBound = 0;
    DevBuf samples, goalSample, nearIdx, nearDist, parent, status;
    size_t parentCap = 0;
    RrtStatus *hStatus = nullptr;   // pinned
    uint64_t iterations = 0, samplesDrawn = 0, edgesChecked = 0;
    int64_t solution = -1;

    ~mplb_rrt() {
        if (hStatus) cudaFreeHost(hStatus);
        if (ctx && scene) scene->release(ctx);
    }
};

static int rrtGrowParents(mplb_rrt *r, size_t need) {
    if (need <= r->parentCap) return MPLB_OK;
    size_t cap = std::max<size_t>(4096, r->parentCap);
    while (cap < need) cap *= 2;
    long long *p = nullptr;
    MPLB_CUDA(cudaStreamSynchronize(r->ctx->stream));
    cudaError_t e = cudaMalloc((void **)&p, cap * 8);
    if (e != cudaSuccess) return fail(MPLB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", cap * 8, cudaGetErrorString(e));
    if (r->parentCap) MPLB_CUDA(cudaMemcpy(p, r->parent.ptr, r->parentCap * 8, cudaMemcpyDeviceToDevice));
    if (r->parent.ptr) cudaFree(r->parent.ptr);
    r->parent.ptr = p;
    r->parent.cap = cap * 8;
    r->parentCap = cap;
    return MPLB_OK;
}

// one iteration; the nearest-neighbour structure is locked by the caller
static int rrtIterate(mplb_rrt *r, double *keys, float *keys32, size_t count) {
    Se3Scene *sc = r->scene;
    CallCtx *c = r->ctx;
    const size_t B = r->batch;
    int rc;
    SampleSe3 sp = r->sampler;
    sp.first = r->iterations * (unsigned long long)B;
    MPLB_CUDA(launchSampleSe3(sp, B, (double *)r->samples.ptr, (uint8_t *)r->goalSample.ptr, c->stream));
    if ((rc = nnSearchDeviceLocked(r->nn, (const double *)r->samples.ptr, B, 1, (long long *)r->nearIdx.ptr,
                                   (double *)r->nearDist.ptr, c->stream)))
        return rc;
    EdgeGather g{};
    g.fromPool = keys;
    g.fromRows = count;
    g.fromIdx = (const long long *)r->nearIdx.ptr;
    g.fromDiv = 1;
    g.toPool = (const double *)r->samples.ptr;
    g.toRows = B;
    g.toIdx = nullptr;
    g.toDiv = 1;
    if ((rc = enqueueIndexed(sc, c, g, B))) return rc;
    RrtSelect sel{};
    sel.samples = (const double *)r->samples.ptr;
    sel.goalSample = (const uint8_t *)r->goalSample.ptr;
    sel.valid = (const uint8_t *)c->out.ptr;
    sel.nearIdx = (const long long *)r->nearIdx.ptr;
    sel.nearDist = (const double *)r->nearDist.ptr;
    sel.keys = keys;
    sel.keys32 = keys32;
    sel.parent = (long long *)r->parent.ptr;
    sel.count = count;
    std::memcpy(sel.goal, r->sampler.goal, sizeof(sel.goal));
    sel.goalRadius = r->goalRadius;
    sel.so3w = sc->so3w;
    sel.l2w = sc->l2w;
    auto select = [&]() -> int {
        MPLB_CUDA(launchRrtSelect(sel, B, (RrtStatus *)r->status.ptr, c->stream));
        MPLB_CUDA(cudaMemcpyAsync(r->hStatus, r->status.ptr, sizeof(RrtStatus), cudaMemcpyDeviceToHost, c->stream));
        return MPLB_OK;
    };
    if ((rc = select())) return rc;
    sc->hLaunches += 6;   // sampler, nearest (3-4 kernels), select
    if ((rc = finishIndexed(sc, c, B, nullptr)) != MPLB_OK) return rc;
    if (c->hCtr->ambiguous) {   // (rare) verdicts may have been replaced after the nodes were chosen: choose again
        if ((rc = select())) return rc;
        MPLB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return MPLB_OK;
}

extern "C" {

int mplb_sample_states_device(mplb_scene *scene, const double *min, const double *max, uint64_t seed, uint64_t first,
                              size_t n, double *d_states, void *stream) {
    if (!scene) return fail(MPLB_ERR_INVALID, "scene is null");
    if (n == 0) return MPLB_OK;
    if (!min || !max || !d_states) return fail(MPLB_ERR_INVALID, "null argument");
    SceneBase *b = reinterpret_cast<SceneBase *>(scene);
    MPLB_CUDA(cudaSetDevice(b->device));
    if (b->stateDim == 7) {
        SampleSe3 p{};
        for (int i = 0; i < 3; ++i) {
            p.lo[i] = min[i];
            p.hi[i] = max[i];
        }
        p.goalBias = 0;
        p.seed = seed;
        p.first = first;
        MPLB_CUDA(launchSampleSe3(p, n, d_states, nullptr, (cudaStream_t)stream));
    } else {
        SampleBox p{};
        p.dim = b->stateDim;
        for (int i = 0; i < p.dim; ++i) {
            p.lo[i] = min[i];
            p.hi[i] = max[i];
        }
        p.seed = seed;
        p.first = first;
        MPLB_CUDA(launchSampleBox(p, n, d_states, (cudaStream_t)stream));
    }
    return MPLB_OK;
}

int mplb_check_edges_indexed(mplb_scene *scene, const double *d_from_pool, size_t from_rows, const int64_t *from_idx,
                             uint32_t from_div, const double *d_to_pool, size_t to_rows, const int64_t *to_idx,
                             uint32_t to_div, size_t n, int idx_on_device, uint8_t *valid, uint64_t *states_checked) {
    if (!scene) return fail(MPLB_ERR_INVALID, "scene is null");
    if (states_checked) *states_checked = 0;
    if (n == 0) return MPLB_OK;
    if (!d_from_pool || !d_to_pool || !valid) return fail(MPLB_ERR_INVALID, "null buffer");
    if (!(n < (1ull << 32))) return fail(MPLB_ERR_INVALID, "too many edges in one call");
    Se3Scene *sc = se3Of(scene);
    if (!sc) return fail(MPLB_ERR_INVALID, "index-named edges are implemented for SE(3) scenes");
    MPLB_CUDA(cudaSetDevice(sc->device));
    CallCtx *c = sc->acquire();
    if (!c) return fail(MPLB_ERR_CUDA, "could not create a call context (stream / pinned memory)");
    CtxGuard guard{sc, c};
    int rc;
    EdgeGather g{};
    g.fromPool = d_from_pool;
    g.toPool = d_to_pool;
    g.fromRows = from_rows;
    g.toRows = to_rows;
    g.fromDiv = std::max(1u, from_div);
    g.toDiv = std::max(1u, to_div);
    g.fromIdx = (const long long *)from_idx;
    g.toIdx = (const long long *)to_idx;
    if (!idx_on_device) {   // the index arrays are the only input that crosses PCIe: 8 B per end point
        if (from_idx) {
            if ((rc = c->in0.reserve(n * 8)) || (rc = hostToDevice(c, c->in0.ptr, from_idx, n * 8, c->stream))) return rc;
            g.fromIdx = (const long long *)c->in0.ptr;
        }
        if (to_idx) {
            if ((rc = c->in1.reserve(n * 8)) || (rc = hostToDevice(c, c->in1.ptr, to_idx, n * 8, c->stream))) return rc;
            g.toIdx = (const long long *)c->in1.ptr;
        }
    }
    if ((rc = enqueueIndexed(sc, c, g, n))) {
        cudaStreamSynchronize(c->stream);
        return rc;
    }
    if ((rc = finishIndexed(sc, c, n, states_checked))) return rc;
    cudaPointerAttributes at{};
    const bool outOnDevice = cudaPointerGetAttributes(&at, valid) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    MPLB_CUDA(cudaMemcpyAsync(valid, c->out.ptr, n, outOnDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    MPLB_CUDA(cudaStreamSynchronize(c->stream));
    return MPLB_OK;
}

int mplb_check_edges_indexed_device(mplb_scene *scene, const double *d_from_pool, size_t from_rows,
                                    const int64_t *d_from_idx, uint32_t from_div, const double *d_to_pool, size_t to_rows,
                                    const int64_t *d_to_idx, uint32_t to_div, size_t n, uint8_t *d_valid, void *stream) {
    if (!scene) return fail(MPLB_ERR_INVALID, "scene is null");
    if (n == 0) return MPLB_OK;
    if (!d_from_pool || !d_to_pool || !d_valid) return fail(MPLB_ERR_INVALID, "null buffer");
    Se3Scene *s = se3Of(scene);
    if (!s) return fail(MPLB_ERR_INVALID, "index-named edges are implemented for SE(3) scenes");
    MPLB_CUDA(cudaSetDevice(s->device));
    {
        std::lock_guard<std::mutex> g(s->mu);
        int rc;
        if ((rc = s->globalDead.reserve(n * 4)) || (rc = s->globalStacks.reserve(edgeStackBytes(s->numSms))) ||
            (rc = s->globalEdgeConst.reserve(edgeGatherBytes(n))))
            return rc;
    }
    EdgeGather g{};
    g.fromPool = d_from_pool;
    g.toPool = d_to_pool;
    g.fromRows = from_rows;
    g.toRows = to_rows;
    g.fromDiv = std::max(1u, from_div);
    g.toDiv = std::max(1u, to_div);
    g.fromIdx = (const long long *)d_from_idx;
    g.toIdx = (const long long *)d_to_idx;
    MPLB_CUDA(launchCheckEdgesIndexed(s->dev, g, n, s->so3w, s->l2w, s->invStep, (unsigned *)s->globalDead.ptr, d_valid,
                                      (DevCounters *)s->globalCtr.ptr, (unsigned long long *)s->globalWork.ptr,
                                      s->globalStacks.ptr, edgeStackBytes(s->numSms), s->globalEdgeConst.ptr, s->numSms,
                                      (cudaStream_t)stream));
    s->hLaunches += 6;
    return MPLB_OK;
}

int mplb_rrt_create(mplb_scene *scene, mplb_nn *nn, const double *start, const double *goal, double goal_radius,
                    const double *min, const double *max, uint64_t seed, size_t batch, double goal_bias, mplb_rrt **out) {
    if (!out) return fail(MPLB_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!scene || !nn || !start || !goal || !min || !max) return fail(MPLB_ERR_INVALID, "null argument");
    if (batch == 0 || batch > (1u << 20)) return fail(MPLB_ERR_INVALID, "batch must be in [1, 2^20]");
    Se3Scene *sc = se3Of(scene);
    if (!sc) return fail(MPLB_ERR_INVALID, "the device-resident loop is implemented for SE(3) scenes");
    int device, metric, dim;
    double so3w, l2w;
    nnInfo(nn, &device, &metric, &dim, &so3w, &l2w);
    if (device != sc->device || metric != 0 || so3w != sc->so3w || l2w != sc->l2w)
        return fail(MPLB_ERR_INVALID, "the nearest-neighbour structure must be an SE3 one on the scene's device with the scene's weights");
    if (mplb_nn_size(nn) != 0) return fail(MPLB_ERR_INVALID, "the nearest-neighbour structure must be empty (it becomes the tree)");
    uint8_t ok = 0;
    int rc = mplb_check_states(scene, start, 1, &ok);
    if (rc != MPLB_OK) return rc;
    if (!ok) return fail(MPLB_ERR_INVALID, "start state is not valid");   // prrt.hpp:109-110
    MPLB_CUDA(cudaSetDevice(sc->device));
    auto r = std::make_unique<mplb_rrt>();
    r->scene = sc;
    r->nn = nn;
    r->batch = batch;
    r->goalRadius = goal_radius;
    for (int i = 0; i < 3; ++i) {
        r->sampler.lo[i] = min[i];
        r->sampler.hi[i] = max[i];
        r->absMaxBound = std::max({r->absMaxBound, std::fabs(min[i]), std::fabs(max[i]), std::fabs(goal[4 + i])});
    }
    std::memcpy(r->sampler.goal, goal, 7 * sizeof(double));
    r->sampler.goalBias = goal_bias;
    r->sampler.seed = seed;
    if ((rc = mplb_nn_insert(nn, start, 1))) return rc;
    r->ctx = sc->acquire();
    if (!r->ctx) return fail(MPLB_ERR_CUDA, "could not create a call context (stream / pinned memory)");
    if ((rc = r->samples.reserve(batch * 56)) || (rc = r->goalSample.reserve(batch)) || (rc = r->nearIdx.reserve(batch * 8)) ||
        (rc = r->nearDist.reserve(batch * 8)) || (rc = r->status.reserve(sizeof(RrtStatus))) || (rc = rrtGrowParents(r.get(), 4 * batch)))
        return rc;
    MPLB_CUDA(cudaMallocHost((void **)&r->hStatus, sizeof(RrtStatus)));
    const long long root = -1;
    MPLB_CUDA(cudaMemcpy(r->parent.ptr, &root, 8, cudaMemcpyHostToDevice));
    *out = r.release();
    return MPLB_OK;
}

void mplb_rrt_destroy(mplb_rrt *rrt) {
    if (!rrt) return;
    cudaSetDevice(rrt->scene->device);
    cudaStreamSynchronize(rrt->ctx->stream);
    // (the parent array was allocated by hand: DevBuf frees it)
    delete rrt;
}

int mplb_rrt_run(mplb_rrt *r, uint64_t max_iterations, double time_limit_s, mplb_rrt_status_t *status) {
    if (!r) return fail(MPLB_ERR_INVALID, "rrt is null");
    MPLB_CUDA(cudaSetDevice(r->scene->device));
    const auto t0 = std::chrono::steady_clock::now();
    int rc = MPLB_OK;
    for (uint64_t it = 0; it < max_iterations && r->solution < 0; ++it) {
        if (time_limit_s > 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > time_limit_s) break;
        double *keys;
        float *keys32;
        size_t count;
        if ((rc = nnLock(r->nn, r->batch, &keys, &keys32, &count))) break;
        if ((rc = rrtGrowParents(r, count + r->batch))) {
            nnUnlock(r->nn, 0, 0);
            break;
        }
        rc = rrtIterate(r, keys, keys32, count);
        if (rc != MPLB_OK) {
            cudaStreamSynchronize(r->ctx->stream);
            nnUnlock(r->nn, 0, 0);
            break;
        }
        const RrtStatus st = *r->hStatus;
        long long goalNode = st.goalNode;
        if (st.goalUncertain) {   // distances within rounding of the goal radius: host libm decides (se3...:115-117)
            std::vector<double> q(7 * st.added);
            cudaMemcpy(q.data(), keys + 7 * count, q.size() * 8, cudaMemcpyDeviceToHost);
            goalNode = -1;
            for (size_t j = 0; j < st.added && goalNode < 0; ++j)
                if (r->scene->distance(r->sampler.goal, &q[7 * j]) <= r->goalRadius) goalNode = (long long)(count + j);
        }
        nnUnlock(r->nn, st.added, r->absMaxBound);
        r->iterations += 1;
        r->samplesDrawn += r->batch;
        r->edgesChecked += st.candidates;
        if (goalNode >= 0) r->solution = goalNode;
    }
    if (status) {
        status->iterations = r->iterations;
        status->samples = r->samplesDrawn;
        status->edges_checked = r->edgesChecked;
        status->nodes = mplb_nn_size(r->nn);
        status->solution = r->solution;
        status->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return rc;
}

int mplb_rrt_tree(mplb_rrt *r, double *nodes, int64_t *parent, size_t capacity, size_t *count) {
    if (!r) return fail(MPLB_ERR_INVALID, "rrt is null");
    MPLB_CUDA(cudaSetDevice(r->scene->device));
    const size_t n = mplb_nn_size(r->nn);
    if (count) *count = n;
    const size_t m = std::min(n, capacity);
    if (m == 0) return MPLB_OK;
    if (nodes) MPLB_CUDA(cudaMemcpy(nodes, mplb_nn_keys_device(r->nn), m * 56, cudaMemcpyDeviceToHost));
    if (parent) MPLB_CUDA(cudaMemcpy(parent, r->parent.ptr, m * 8, cudaMemcpyDeviceToHost));
    return MPLB_OK;
}

int mplb_rrt_samples(mplb_rrt *r, uint64_t iteration, double *states, uint8_t *is_goal_sample) {
    if (!r || !states) return fail(MPLB_ERR_INVALID, "null argument");
    MPLB_CUDA(cudaSetDevice(r->scene->device));
    DevBuf s, g;
    int rc;
    if ((rc = s.reserve(r->batch * 56)) || (rc = g.reserve(r->batch))) return rc;
    SampleSe3 sp = r->sampler;
    sp.first = iteration * (unsigned long long)r->batch;
    MPLB_CUDA(launchSampleSe3(sp, r->batch, (double *)s.ptr, (uint8_t *)g.ptr, nullptr));
    MPLB_CUDA(cudaMemcpy(states, s.ptr, r->batch * 56, cudaMemcpyDeviceToHost));
    if (is_goal_sample) MPLB_CUDA(cudaMemcpy(is_goal_sample, g.ptr, r->batch, cudaMemcpyDeviceToHost));
    return MPLB_OK;
}

}  // extern "C"
