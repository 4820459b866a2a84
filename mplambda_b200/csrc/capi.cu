// C-ABI of libmplb.so (declared in include/mplb.h): scene lifetime, host<->device staging,
// per-call streams.  No CPU fallback: without a CUDA device every compute entry point fails.
#include "../../include/mplb.h"
#include "mplb_host.hpp"
#include "se3_device.cuh"
#include "scene_impl.cuh"

#include <cuda.h>   // types of the stream memory operation only: the entry point is resolved at run time

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <memory>
#include <thread>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace mplb {

thread_local std::string gLastError;

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    gLastError = buf;
    return code;
}

// ---- device buffer that only grows
int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return MPLB_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = std::max(bytes, (size_t)4096);
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) return fail(MPLB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    cap = want;
    return MPLB_OK;
}
DevBuf::~DevBuf() {
    if (ptr) cudaFree(ptr);
}

static inline void cpuRelax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}

constexpr size_t kStageBytes = 160 * 1024;                 // inputs of a small call
constexpr size_t kCtlHeader = ((sizeof(DevCounters) + 127) / 128) * 128 + 128;   // counters, then 16 work counters
constexpr size_t kCtlTail = 8 * 1024;                      // verdicts (1 B/state) or dead flags (4 B/edge)
constexpr size_t kSmallStates = 2048, kSmallEdges = 1024;
static_assert(kSmallStates * 56 <= kStageBytes && kSmallEdges * (116 + 40) + 8 <= kStageBytes, "staging too small");
static_assert(kSmallStates <= kCtlTail && kSmallEdges * 4 <= kCtlTail, "control block too small");

constexpr size_t kBounceBytes = 8u << 20;

int CallCtx::reserveRes(int i, size_t bytes) {
    if (bytes <= hResCap[i]) return MPLB_OK;
    if (hRes[i]) cudaFreeHost(hRes[i]);
    hRes[i] = nullptr;
    hResCap[i] = 0;
    if (cudaMallocHost((void **)&hRes[i], bytes) != cudaSuccess) return fail(MPLB_ERR_NOMEM, "out of pinned host memory");
    hResCap[i] = bytes;
    return MPLB_OK;
}

CallCtx::~CallCtx() {
    if (graphState) cudaGraphExecDestroy(graphState);
    if (graphEdge) cudaGraphExecDestroy(graphEdge);
    for (int i = 0; i < 2; ++i) {
        if (hRes[i]) cudaFreeHost(hRes[i]);
        if (hBounce[i]) cudaFreeHost(hBounce[i]);
        if (bounceDone[i]) cudaEventDestroy(bounceDone[i]);
    }
    if (hStage) cudaFreeHost(hStage);
    if (hCtl) cudaFreeHost(hCtl);
    if (hCtr) cudaFreeHost(hCtr);
    if (hWater) cudaFreeHost(hWater);
    for (auto e : events) cudaEventDestroy(e);
    if (copyStream) cudaStreamDestroy(copyStream);
    if (copyStream2) cudaStreamDestroy(copyStream2);
    if (stream) cudaStreamDestroy(stream);
}

CallCtx *SceneBase::acquire() {
    {
        std::lock_guard<std::mutex> g(mu);
        if (!freeCtx.empty()) {
            CallCtx *c = freeCtx.back();
            freeCtx.pop_back();
            return c;
        }
    }
    auto *c = new CallCtx();
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copyStream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost((void **)&c->hCtr, sizeof(DevCounters)) != cudaSuccess ||
        cudaMallocHost((void **)&c->hWater, 64 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMallocHost((void **)&c->hStage, kStageBytes) != cudaSuccess ||
        cudaMallocHost((void **)&c->hCtl, kCtlHeader + kCtlTail) != cudaSuccess ||
        c->stage.reserve(kStageBytes) != MPLB_OK || c->ctl.reserve(kCtlHeader + kCtlTail) != MPLB_OK ||
        c->ctr.reserve(sizeof(DevCounters)) != MPLB_OK || c->work.reserve(128) != MPLB_OK) {
        delete c;
        return nullptr;
    }
    static const bool noDirect = std::getenv("MPLB_NO_DIRECT_RETURN") != nullptr;   // diagnostic
    void *ds = nullptr, *dc = nullptr;
    if (!noDirect && cudaHostGetDevicePointer(&ds, c->hStage, 0) == cudaSuccess &&
        cudaHostGetDevicePointer(&dc, c->hCtl, 0) == cudaSuccess) {
        c->dStageHost = (unsigned char *)ds;
        c->dCtlHost = (unsigned char *)dc;
    } else {
        cudaGetLastError();
    }
    std::memset(c->hCtl, 0, kCtlHeader + kCtlTail);
    return c;
}

// Waits for a direct-return kernel to ring (HostReturn): the calling thread spins on the flag in pinned memory -- the
// whole call is tens of microseconds, a blocking synchronisation alone costs that much -- and looks at the stream now
// and then, so that a failed launch or a device fault ends the wait with an error instead of a hang.
static int awaitRing(CallCtx *c, unsigned seq) {
    volatile unsigned *flag = (volatile unsigned *)(c->hCtl + kCtlHeader - 64);
    for (unsigned spins = 1;; ++spins) {
        if (*flag == seq) break;
        if ((spins & 0xfffu) == 0) {
            const cudaError_t q = cudaStreamQuery(c->stream);
            if (q == cudaSuccess) {
                if (*flag == seq) break;
                c->ctlDirty = true;
                return fail(MPLB_ERR_CUDA, "direct-return kernel finished without ringing the host");
            }
            if (q != cudaErrorNotReady) {
                c->ctlDirty = true;
                return fail(MPLB_ERR_CUDA, "device work failed: %s", cudaGetErrorString(q));
            }
        }
        cpuRelax();
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return MPLB_OK;
}
void SceneBase::release(CallCtx *c) {
    std::lock_guard<std::mutex> g(mu);
    freeCtx.push_back(c);
}
SceneBase::~SceneBase() {
    for (auto *c : freeCtx) delete c;
}

// One scalar call through the combiner.  The first caller of a batch leads it: it takes the batch to the device as
// soon as fewer than `maxInflight` batches of this kind are there (at once when the scene is idle, so a single-threaded
// planner pays nothing), and whoever arrives in the meantime joins; the others wait for the leader's launch, spinning
// first (the wait is tens of microseconds; a futex wake-up alone costs that much).  A failed launch is not shared:
// every caller of that batch repeats its own call, so one thread's bad state never fails another's.
int SceneBase::combined(int kind, const double *a, const double *b, uint8_t *valid) {
    static const int maxInflight = [] { const char *e = std::getenv("MPLB_COALESCE_INFLIGHT"); return e ? std::max(1, std::atoi(e)) : 2; }();
    Combiner &cb = comb[kind];
    const size_t w = (size_t)stateDim;
    std::shared_ptr<Combiner::Batch> bt;
    int i;
    {
        std::lock_guard<std::mutex> g(cb.mu);
        if (!cb.open) {
            if (!cb.spare.empty()) {
                cb.open = std::move(cb.spare.back());
                cb.spare.pop_back();
            } else {
                cb.open = std::make_shared<Combiner::Batch>();
                cb.open->a.resize(Combiner::kMax * w);
                cb.open->b.resize(Combiner::kMax * w);
            }
            cb.open->n = 0;
            cb.open->filled.store(0, std::memory_order_relaxed);
            cb.open->done.store(0, std::memory_order_relaxed);
            cb.open->left.store(0, std::memory_order_relaxed);
        }
        bt = cb.open;
        i = bt->n++;
        if (bt->n == Combiner::kMax) cb.open.reset();   // full: newcomers start the next one
    }
    std::memcpy(&bt->a[i * w], a, w * sizeof(double));
    if (kind) std::memcpy(&bt->b[i * w], b, w * sizeof(double));
    bt->filled.fetch_add(1, std::memory_order_release);
    if (i == 0) {
        while (cb.inflight.load(std::memory_order_acquire) >= maxInflight) {
            bool full;
            {
                std::lock_guard<std::mutex> g(cb.mu);
                full = bt->n == Combiner::kMax;
            }
            if (full) break;
            cpuRelax();
        }
        int n;
        {
            std::lock_guard<std::mutex> g(cb.mu);
            if (cb.open == bt) cb.open.reset();   // closed
            n = bt->n;
        }
        cb.inflight.fetch_add(1, std::memory_order_acq_rel);
        while (bt->filled.load(std::memory_order_acquire) < n) cpuRelax();   // (a joiner still copying its state in)
        const int rc = kind ? checkEdges(bt->a.data(), bt->b.data(), (size_t)n, bt->out, nullptr)
                            : checkStates(bt->a.data(), (size_t)n, bt->out);
        cb.inflight.fetch_sub(1, std::memory_order_acq_rel);
        if (n > 1) hCoalesced += (uint64_t)n;
        bt->done.store(rc == MPLB_OK ? 1 : 2, std::memory_order_release);
    } else {
        for (unsigned spins = 0; bt->done.load(std::memory_order_acquire) == 0; ++spins) {
            if (spins < 4096) cpuRelax();
            else std::this_thread::yield();
        }
    }
    const bool ok = bt->done.load(std::memory_order_acquire) == 1;
    const uint8_t v = bt->out[i];
    // bt->n is final (the batch was closed before its launch); the last caller out hands the batch back for reuse
    if (bt->left.fetch_add(1, std::memory_order_acq_rel) + 1 == bt->n) {
        std::lock_guard<std::mutex> g(cb.mu);
        if (cb.spare.size() < 8) cb.spare.push_back(std::move(bt));
    }
    if (!ok) return kind ? checkEdges(a, b, 1, valid, nullptr) : checkStates(a, 1, valid);
    *valid = v;
    return MPLB_OK;
}

int bitsFor(size_t n) {
    int b = 1;
    while (((size_t)1 << b) < n) ++b;
    return b;
}

int selectDevice(int device, int *numSms) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(MPLB_ERR_CUDA, "no CUDA device available (%s); libmplb has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(MPLB_ERR_INVALID, "device %d out of range [0,%d)", device, count);
    MPLB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MPLB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(MPLB_ERR_CUDA, "device %d is sm_%d%d; libmplb is built for sm_100a only", device, prop.major, prop.minor);
    *numSms = prop.multiProcessorCount;
    return MPLB_OK;
}

static int se3CheckStates(Se3Scene *sc, const double *states, size_t n, uint8_t *valid);
static int se3CheckEdges(Se3Scene *sc, const double *from, const double *to, size_t n, uint8_t *valid,
                         uint64_t *statesChecked);

static int createSe3(const double *envTris, size_t nEnv, const double *robotTris, size_t nRobot, double so3w,
                     double l2w, double res, int device, mplb_scene **out) {
    if (!out) return fail(MPLB_ERR_INVALID, "out is null");
    *out = nullptr;
    if ((nEnv && !envTris) || (nRobot && !robotTris)) return fail(MPLB_ERR_INVALID, "null triangle array");
    if (!(res > 0)) return fail(MPLB_ERR_INVALID, "check_resolution must be > 0");
    for (size_t i = 0; i < 9 * nEnv; ++i)
        if (!std::isfinite(envTris[i])) return fail(MPLB_ERR_INVALID, "non-finite environment vertex");
    for (size_t i = 0; i < 9 * nRobot; ++i)
        if (!std::isfinite(robotTris[i])) return fail(MPLB_ERR_INVALID, "non-finite robot vertex");
    int numSms = 0;
    int rc = selectDevice(device, &numSms);
    if (rc != MPLB_OK) return rc;

    FlatBvh env = buildBvhCached(envTris, nEnv), robot = buildBvhCached(robotTris, nRobot);
    auto sc = std::make_unique<Se3Scene>();
    sc->device = device;
    sc->numSms = numSms;
    sc->so3w = so3w;
    sc->l2w = l2w;
    sc->invStep = 1 / res;
    sc->nEnvTris = nEnv;
    sc->nRobotTris = nRobot;
    if ((rc = upload(sc->robotNodes, robot.nodes)) || (rc = upload(sc->envNodes, env.nodes)) ||
        (rc = upload(sc->robotTris, robot.tris32)) || (rc = upload(sc->envTris, env.tris32)) ||
        (rc = upload(sc->robotTris64, robot.tris64)) || (rc = upload(sc->envTris64, env.tris64)))
        return rc;
    if ((rc = sc->globalCtr.reserve(sizeof(DevCounters))) || (rc = sc->globalWork.reserve(128))) return rc;
    MPLB_CUDA(cudaMemset(sc->globalCtr.ptr, 0, sizeof(DevCounters)));
    SceneDev &d = sc->dev;
    d.robotNodes = (const float4 *)sc->robotNodes.ptr;
    d.envNodes = (const float4 *)sc->envNodes.ptr;
    d.robotTris = (const float4 *)sc->robotTris.ptr;
    d.envTris = (const float4 *)sc->envTris.ptr;
    d.robotTris64 = (const double *)sc->robotTris64.ptr;
    d.envTris64 = (const double *)sc->envTris64.ptr;
    d.nRobotNodes = (int)robot.nodes.size();
    d.nEnvNodes = (int)env.nodes.size();
    d.haveGeometry = nEnv > 0 && nRobot > 0;
    d.bothRootsLeaf = nEnv == 1 && nRobot == 1;
    d.bitsA = bitsFor(std::max<size_t>(2, robot.nodes.size()));
    d.bitsB = bitsFor(std::max<size_t>(2, env.nodes.size()));
    d.wide = (6 + d.bitsA + d.bitsB) > 32;
    if (d.bitsA + d.bitsB > 32) return fail(MPLB_ERR_INVALID, "meshes too large for the 32-bit node pair encoding");
    // se3_kernels.cu: a lane's depth-first LIFO holds at most depthA + depthB + 2 pairs (kStackDepth = 64)
    if (robot.depth + env.depth > 60)
        return fail(MPLB_ERR_INVALID, "hierarchy too deep for the device pair stack (%d + %d levels)", robot.depth,
                    env.depth);
    d.envRadius = std::nextafterf((float)env.radius, INFINITY);
    d.robotRadius = std::nextafterf((float)robot.radius, INFINITY);
    d.stackLevels = (robot.depth + env.depth + 2 + 3) & ~3;
    {
        // entry layout of se3_kernels.cu: robotIdx | envIdx << bitsA | descendRobot << 31, the index on
        // the descended side being that node's first child
        auto linkOf = [](const FlatBvh &b) { int l = -1; if (!b.nodes.empty()) std::memcpy(&l, &b.nodes[0].q[3][3], 4); return l; };
        auto sizeOf = [](const FlatBvh &b) { if (b.nodes.empty()) return 0.f; const float *e = b.nodes[0].q[3]; return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; };
        const int lr = linkOf(robot), le = linkOf(env);
        const bool descendRobot = le < 0 || (lr >= 0 && sizeOf(robot) > sizeOf(env));
        d.rootEntry = (lr < 0 && le < 0) ? 0u
                      : descendRobot ? ((unsigned)lr | 0u << d.bitsA | 0x80000000u) : (0u | (unsigned)le << d.bitsA);
    }
    if (d.bitsA + d.bitsB > 31) return fail(MPLB_ERR_INVALID, "meshes too large for the 31-bit node pair encoding");
    d.triBatchMin = 32;
    d.triEagerLanes = 16;
    if (const char *e = std::getenv("MPLB_TRI_EAGER")) d.triEagerLanes = std::atoi(e);
    if (const char *e = std::getenv("MPLB_TRI_BATCH")) d.triBatchMin = std::max(1, std::min(32, std::atoi(e)));
    sc->envNodeCount = env.nodes.size();
    sc->robotNodeCount = robot.nodes.size();
    {
        // first-use costs (kernel load, call context, interval scratch: ~0.1 s) belong to construction, not
        // to the planner's first isValid
        const double q[7] = {0, 0, 0, 1, 0, 0, 0};
        uint8_t v;
        if ((rc = se3CheckStates(sc.get(), q, 1, &v)) || (rc = se3CheckEdges(sc.get(), q, q, 1, &v, nullptr))) return rc;
        sc->hChecks = sc->hBv = sc->hTri = sc->hExact = sc->hLaunches = 0;
    }
    *out = reinterpret_cast<mplb_scene *>(static_cast<SceneBase *>(sc.release()));
    return MPLB_OK;
}

bool isPinned(const void *p) {
    cudaPointerAttributes at{};
    const bool pinned = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();   // (older drivers report plain malloc memory as an error)
    return pinned;
}

// Host -> device copy of a caller's buffer on `stream`.  Pinned memory goes straight to the copy engine; pageable
// memory is copied by all cores into two pinned bounce buffers, piece i + 1 being filled while piece i is on the wire
// (the buffers alternate across calls too: c->bounceNext).
int hostToDevice(CallCtx *c, void *dst, const void *src, size_t bytes, cudaStream_t stream) {
    static const bool off = std::getenv("MPLB_NO_BOUNCE") != nullptr;   // diagnostic: let the driver stage pageable memory
    if (bytes < (1u << 20) || off || isPinned(src)) {
        MPLB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return MPLB_OK;
    }
    for (int i = 0; i < 2; ++i)
        if (!c->hBounce[i]) {
            MPLB_CUDA(cudaMallocHost((void **)&c->hBounce[i], kBounceBytes));
            MPLB_CUDA(cudaEventCreateWithFlags(&c->bounceDone[i], cudaEventDisableTiming));
        }
    for (size_t at = 0; at < bytes; at += kBounceBytes) {
        const int b = c->bounceNext;
        c->bounceNext ^= 1;
        const size_t len = std::min(kBounceBytes, bytes - at);
        MPLB_CUDA(cudaEventSynchronize(c->bounceDone[b]));   // the buffer's previous piece has left
        const unsigned char *s = (const unsigned char *)src + at;
        unsigned char *d = c->hBounce[b];
        const long long parts = (long long)((len + (1u << 20) - 1) >> 20);
#pragma omp parallel for schedule(static)
        for (long long p = 0; p < parts; ++p) {
            const size_t lo = (size_t)p << 20, n = std::min((size_t)1 << 20, len - lo);
            std::memcpy(d + lo, s + lo, n);
        }
        MPLB_CUDA(cudaMemcpyAsync((unsigned char *)dst + at, d, len, cudaMemcpyHostToDevice, stream));
        MPLB_CUDA(cudaEventRecord(c->bounceDone[b], stream));
    }
    return MPLB_OK;
}

// Runs `enqueue` (asynchronous work on c->stream, fixed addresses and sizes) through a cached CUDA graph: replayed if
// there is one for these buffers, captured + instantiated on the context's second such call, enqueued directly
// otherwise (and whenever capturing fails: graphs are an optimisation, never a requirement).
template <class Enqueue>
static int runScalar(CallCtx *c, cudaGraphExec_t &exec, void **key, const void *const *now, int nKeys, unsigned &calls,
                     Enqueue enqueue) {
    static const bool off = std::getenv("MPLB_NO_GRAPHS") != nullptr;
    bool same = exec != nullptr;
    for (int i = 0; i < nKeys && same; ++i) same = key[i] == now[i];
    if (same) {
        MPLB_CUDA(cudaGraphLaunch(exec, c->stream));
        return MPLB_OK;
    }
    if (off || c->graphsOff || ++calls < 2) return enqueue();
    if (exec) {
        cudaGraphExecDestroy(exec);
        exec = nullptr;
    }
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        c->graphsOff = true;
        return enqueue();
    }
    const int rc = enqueue();
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc != MPLB_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        exec = nullptr;
        c->graphsOff = true;
        return enqueue();   // nothing has run yet: do it the plain way
    }
    cudaGraphDestroy(graph);
    for (int i = 0; i < nKeys; ++i) key[i] = const_cast<void *>(now[i]);
    MPLB_CUDA(cudaGraphLaunch(exec, c->stream));
    return MPLB_OK;
}

static int account(SceneBase *sc, const DevCounters &k, size_t checks, uint64_t *statesOut) {
    if (k.overflow) return fail(MPLB_ERR_CUDA, "device traversal bookkeeping error (pair stack / slot accounting)");
    sc->hBv += k.bvTests;
    sc->hTri += k.triTests;
    sc->hExact += k.exactTests;
    sc->hChecks += checks ? checks : k.states;
    if (statesOut) *statesOut = k.states;
    return MPLB_OK;
}

static int finishCall(SceneBase *sc, CallCtx *c, size_t checks, uint64_t *statesOut) {
    cudaError_t e = cudaMemcpyAsync(c->hCtr, c->ctr.ptr, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(MPLB_ERR_CUDA, "device work failed: %s", cudaGetErrorString(e));
#ifdef MPLB_WARP_TIMES
    {
        const DevCounters &k = *c->hCtr;
        const double t0 = (double)~k.negStart;
        std::fprintf(stderr, "[mplb warp times] first warp out of new work at %.1f us, first warp done %.1f us, last warp done %.1f us\n",
                     ((double)~k.negFirstDry - t0) * 1e-3, ((double)~k.negFirstDone - t0) * 1e-3, ((double)k.lastDone - t0) * 1e-3);
        std::vector<int> ord;
        for (int w = 0; w < 4096; ++w)
            if (k.warpDone[w]) ord.push_back(w);
        std::sort(ord.begin(), ord.end(), [&](int a, int b) { return k.warpDone[a] > k.warpDone[b]; });
        for (size_t i = 0; i < ord.size() && i < 6; ++i)
            std::fprintf(stderr, "    warp %4d done %.1f us: states %u bv %u tri %u\n", ord[i], ((double)k.warpDone[ord[i]] - t0) * 1e-3,
                         k.warpStates[ord[i]], k.warpBv[ord[i]], k.warpTri[ord[i]]);
        if (!ord.empty()) {
            std::fprintf(stderr, "    warps done (us) by percentile, %zu warps:", ord.size());
            for (int pc : {0, 10, 25, 50, 75, 90, 95, 99, 100})
                std::fprintf(stderr, " p%d %.0f", pc, ((double)k.warpDone[ord[(ord.size() - 1) * (100 - pc) / 100]] - t0) * 1e-3);
            std::vector<double> dry, gap;
            for (int w : ord)
                if (k.warpDry[w] != ~0ull && k.warpDry[w]) {
                    dry.push_back(((double)k.warpDry[w] - t0) * 1e-3);
                    gap.push_back(((double)k.warpDone[w] - (double)k.warpDry[w]) * 1e-3);
                }
            std::sort(dry.begin(), dry.end());
            std::sort(gap.begin(), gap.end());
            if (!dry.empty()) {
                std::fprintf(stderr, "\n    warps dry (us):");
                for (int pc : {0, 10, 50, 90, 100}) std::fprintf(stderr, " p%d %.0f", pc, dry[(dry.size() - 1) * pc / 100]);
                std::fprintf(stderr, "\n    dry -> done (us):");
                for (int pc : {0, 10, 25, 50, 75, 90, 99, 100}) std::fprintf(stderr, " p%d %.0f", pc, gap[(gap.size() - 1) * pc / 100]);
            }
            std::fprintf(stderr, "\n");
        }
    }
#endif
    return account(sc, *c->hCtr, checks, statesOut);
}

// ---- stream memory operations: cuStreamWriteValue32 resolved at run time (no link-time dependency on libcuda)
typedef CUresult (*WriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static WriteValue32Fn streamWriteValue32() {
    static const WriteValue32Fn fn = []() -> WriteValue32Fn {
        if (std::getenv("MPLB_NO_STREAM_MEMOPS")) return nullptr;   // diagnostic: exercise the chunked fallback
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (WriteValue32Fn)p;
    }();
    return fn;
}

// Large host batches stream in under the running kernel.  The batch is copied in pieces on the context's copy stream
// and, behind each piece on that stream, a stream memory operation (cuStreamWriteValue32: ordered after the copy, with
// a system-wide fence in front of the write) advances the call's watermark word.  The kernel is launched first on the
// compute stream, hands states out in index order and picks a claim up once the watermark covers it (acquire load).
// The input is 56 B/state and the kernel needs ~1 ns/state, so the call costs about one host->device copy of the batch.
// Without stream memory operations the watermark travels as a 4-byte copy on the same stream (stream order: it starts
// after the piece's copy has completed).  When a kernel reports that its input did not arrive in time (a profiler
// serialising the kernel and the copies) the batch is repeated piece by piece copy-then-kernel.
static bool streamBatch(size_t n) {
    static const long forced = [] {
        const char *e = std::getenv("MPLB_STREAM_MIN");
        return e ? std::atol(e) : 0L;
    }();
    return n >= (size_t)(forced > 0 ? forced : 65536) && n < ((size_t)1 << 32);
}

// Pieces a streamed batch is copied in (states per piece, multiples of 1024).  Measured on 2^20 Alpha poses: the copy is
// the longer leg (1.06 ms in one piece, 1.11 ms in eight with the watermark writes on a side stream, 1.35 ms in 64); the
// kernel finishes ~0.23 ms after the last piece whatever the piece size or order (halving pieces at the end, small
// pieces in front, 64 K or 256 K pieces: all within 2 %) -- that is the time a warp needs to drain the states it
// holds, not the size of the last piece.
constexpr size_t kMaxPieces = 128;   // (CallCtx::hWater holds 128 words)
static std::vector<size_t> streamPieces(size_t n) {
    static const long forced = [] {
        const char *e = std::getenv("MPLB_STREAM_CHUNK");
        return e ? std::atol(e) : 0L;
    }();
    auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
    const size_t piece = std::max(up(forced > 0 ? (size_t)forced : 131072), up(n / (kMaxPieces - 1)));
    std::vector<size_t> pieces;
    for (size_t left = n; left > 0; left -= pieces.back()) pieces.push_back(std::min(piece, left));
    return pieces;
}

// both streams idle before the context goes back to the pool, whatever path the call leaves by
struct StreamsQuiet {
    CallCtx *c;
    ~StreamsQuiet() {
        cudaStreamSynchronize(c->copyStream);
        cudaStreamSynchronize(c->copyStream2);
        cudaStreamSynchronize(c->stream);
    }
};

// copy-then-kernel in a few pieces: piece i + 1 is on the wire while piece i is being checked
static int se3StatesPipelined(Se3Scene *sc, CallCtx *c, const double *states, size_t n, double *dIn, uint8_t *dOut) {
    const size_t pieces = n >= 4 * 16384 ? 4 : 1, per = (n + pieces - 1) / pieces;
    while (c->events.size() < pieces) {
        cudaEvent_t e;
        MPLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->events.push_back(e);
    }
    int rc;
    for (size_t p = 0, lo = 0; lo < n; ++p, lo += per) {
        const size_t cnt = std::min(per, n - lo);
        if ((rc = hostToDevice(c, dIn + 7 * lo, states + 7 * lo, cnt * 7 * sizeof(double), c->copyStream))) return rc;
        MPLB_CUDA(cudaEventRecord(c->events[p], c->copyStream));
        MPLB_CUDA(cudaStreamWaitEvent(c->stream, c->events[p], 0));
        MPLB_CUDA(launchCheckStates(sc->dev, dIn + 7 * lo, cnt, dOut + lo, (DevCounters *)c->ctr.ptr, (unsigned *)c->work.ptr,
                                    sc->numSms, c->stream));
        sc->hLaunches += 1;
    }
    return MPLB_OK;
}

static int se3CheckStates(Se3Scene *sc, const double *states, size_t n, uint8_t *valid) {
    MPLB_CUDA(cudaSetDevice(sc->device));
    CallCtx *c = sc->acquire();
    if (!c) return fail(MPLB_ERR_CUDA, "could not create a call context (stream / pinned memory)");
    CtxGuard g{sc, c};
    int rc;
    const size_t bytes = n * 7 * sizeof(double);
    if (n <= kSmallStates) {
        std::memcpy(c->hStage, states, bytes);
        unsigned char *ctl = (unsigned char *)c->ctl.ptr;
        if (c->dStageHost) {
            // direct return: one launch, no copy engine, no stream synchronisation (se3_device.cuh: HostReturn)
            if (c->ctlDirty) {
                MPLB_CUDA(cudaMemsetAsync(ctl, 0, kCtlHeader + kCtlTail, c->stream));
                c->ctlDirty = false;
            }
            StreamedInput in;
            in.ret.hostCtr = (DevCounters *)c->dCtlHost;
            in.ret.flag = (unsigned *)(c->dCtlHost + kCtlHeader - 64);
            in.ret.seq = ++c->seq;
            in.ret.clearWords = (unsigned)((kCtlHeader - sizeof(DevCounters)) / 4);
            MPLB_CUDA(launchCheckStates(sc->dev, (const double *)c->dStageHost, n, c->dCtlHost + kCtlHeader, (DevCounters *)ctl,
                                        (unsigned *)(ctl + kCtlHeader - 128), sc->numSms, c->stream, in, true));
            sc->hLaunches += 1;
            if ((rc = awaitRing(c, in.ret.seq))) return rc;
            std::memcpy(valid, c->hCtl + kCtlHeader, n);
            return account(sc, *(const DevCounters *)c->hCtl, n, nullptr);
        }
        c->ctlDirty = true;
        auto enqueue = [&]() -> int {
            MPLB_CUDA(cudaMemsetAsync(ctl, 0, kCtlHeader, c->stream));
            MPLB_CUDA(cudaMemcpyAsync(c->stage.ptr, c->hStage, bytes, cudaMemcpyHostToDevice, c->stream));
            MPLB_CUDA(launchCheckStates(sc->dev, (const double *)c->stage.ptr, n, ctl + kCtlHeader, (DevCounters *)ctl,
                                        (unsigned *)(ctl + kCtlHeader - 128), sc->numSms, c->stream, StreamedInput(), true));
            MPLB_CUDA(cudaMemcpyAsync(c->hCtl, ctl, kCtlHeader + n, cudaMemcpyDeviceToHost, c->stream));
            return MPLB_OK;
        };
        if (n == 1) {
            const void *now[2] = {c->stage.ptr, c->ctl.ptr};
            if ((rc = runScalar(c, c->graphState, c->graphStateKey, now, 2, c->scalarStates, enqueue))) return rc;
        } else if ((rc = enqueue())) {
            return rc;
        }
        sc->hLaunches += 1;
        MPLB_CUDA(cudaStreamSynchronize(c->stream));
        std::memcpy(valid, c->hCtl + kCtlHeader, n);
        return account(sc, *(const DevCounters *)c->hCtl, n, nullptr);
    }
    if ((rc = c->in0.reserve(bytes)) || (rc = c->out.reserve(n))) return rc;
    double *dIn = (double *)c->in0.ptr;
    uint8_t *dOut = (uint8_t *)c->out.ptr;
    // mode of a large batch: 2 = the kernel pulls the batch out of mapped page-locked host memory itself, 1 = chunked
    // copies with arrival flags under the running kernel, 0 = copy-then-kernel in pieces
    static const int forcedMode = [] { const char *e = std::getenv("MPLB_STREAM_MODE"); return e ? std::atoi(e) : -1; }();
    const WriteValue32Fn writeValue = streamWriteValue32();
    void *mapped = nullptr;
    if (streamBatch(n) && forcedMode == 2 && isPinned(states)) {
        if (cudaHostGetDevicePointer(&mapped, const_cast<double *>(states), 0) != cudaSuccess) {
            cudaGetLastError();
            mapped = nullptr;
        }
    }
    // Streaming (the kernel picks the batch up while it is still arriving) is for page-locked inputs only: there the
    // host thread merely enqueues asynchronous copies.  A pageable input goes through the bounce buffers, filled by this
    // thread's OpenMP team -- with several callers at once that team can be descheduled for longer than a resident
    // kernel should ever wait for its input, so those batches go copy-then-kernel in pieces.
    const bool streamed = !mapped && streamBatch(n) && forcedMode != 0 && isPinned(states) && !sc->dev.minMargin &&
                          sc->streamStallsInARow.load(std::memory_order_relaxed) < 2;
    // MPLB_TIMELINE=1: print where a streamed call's time goes (diagnostic)
    static const bool timeline = std::getenv("MPLB_TIMELINE") != nullptr;
    static thread_local cudaEvent_t tl[5] = {};
    const auto h0 = std::chrono::steady_clock::now();
    if (timeline && !tl[0])
        for (auto &e : tl) cudaEventCreate(&e);
    StreamsQuiet quiet{c};
    MPLB_CUDA(cudaMemsetAsync(c->ctr.ptr, 0, sizeof(DevCounters), c->stream));
    if (mapped) {
        StreamedInput in;
        in.host = (const double *)mapped;
        MPLB_CUDA(launchCheckStates(sc->dev, dIn, n, dOut, (DevCounters *)c->ctr.ptr, (unsigned *)c->work.ptr, sc->numSms,
                                    c->stream, in));
        sc->hLaunches += 1;
    } else if (!streamed) {
        if ((rc = se3StatesPipelined(sc, c, states, n, dIn, dOut))) return rc;
    } else {
        if (!c->flags.ptr) {
            if ((rc = c->flags.reserve(128))) return rc;
            MPLB_CUDA(cudaMemsetAsync(c->flags.ptr, 0, 128, c->stream));
        }
        c->epoch = c->epoch % 1000u + 1u;   // never 0 (the word starts at 0), never the previous call's
        StreamedInput in;
        in.water = (const unsigned *)c->flags.ptr;
        in.epoch = c->epoch;
        // MPLB_STREAM_PRECOPY (diagnostic): the streamed kernel on a batch that is already there
        static const bool precopy = std::getenv("MPLB_STREAM_PRECOPY") != nullptr;
        auto launch = [&]() -> int {
            if (timeline) cudaEventRecord(tl[1], c->stream);
            MPLB_CUDA(launchCheckStates(sc->dev, dIn, n, dOut, (DevCounters *)c->ctr.ptr, (unsigned *)c->work.ptr, sc->numSms,
                                        c->stream, in));
            sc->hLaunches += 1;
            if (timeline) cudaEventRecord(tl[3], c->stream);
            return MPLB_OK;
        };
        if (timeline) cudaEventRecord(tl[0], c->copyStream);
        // The copies are enqueued before the kernel is launched: the copy engine starts at once and the kernel joins a
        // few tens of microseconds later (the copy is the longer leg by far), and a host thread that loses the processor
        // in between leaves nothing on the device waiting for it.
        // MPLB_STREAM_FLAGCOPY (diagnostic): the watermark goes over as a 4-byte copy from page-locked memory instead
        static const bool flagCopy = std::getenv("MPLB_STREAM_FLAGCOPY") != nullptr;
        const std::vector<size_t> pieces = streamPieces(n);
        while (c->events.size() < pieces.size()) {
            cudaEvent_t e;
            MPLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->events.push_back(e);
        }
        // The pieces go back to back on the copy stream with an event behind each; the watermark writes wait for those
        // events on a second stream, so the copy engine never waits for a write between two pieces.
        size_t lo = 0;
        for (size_t k = 0; k < pieces.size(); ++k) {
            const size_t cnt = pieces[k];
            if ((rc = hostToDevice(c, dIn + 7 * lo, states + 7 * lo, cnt * 7 * sizeof(double), c->copyStream))) return rc;
            lo += cnt;
            MPLB_CUDA(cudaEventRecord(c->events[k], c->copyStream));
            MPLB_CUDA(cudaStreamWaitEvent(c->copyStream2, c->events[k], 0));
            cudaStream_t ws = c->copyStream2;
            const unsigned mark = c->epoch << kWaterShift | (unsigned)((lo + 1023) >> kWaterUnitBits);
            if (flagCopy || !writeValue) {
                unsigned *h = (unsigned *)c->hWater + k;
                *h = mark;
                MPLB_CUDA(cudaMemcpyAsync(c->flags.ptr, h, sizeof(unsigned), cudaMemcpyHostToDevice, ws));
            } else {
                const CUresult wr = writeValue((CUstream)ws, (CUdeviceptr)c->flags.ptr, mark, 0);
                if (wr != CUDA_SUCCESS) return fail(MPLB_ERR_CUDA, "cuStreamWriteValue32 failed (%d)", (int)wr);
            }
        }
        if (timeline) cudaEventRecord(tl[2], c->copyStream);
        if (precopy) MPLB_CUDA(cudaStreamSynchronize(c->copyStream));
        if ((rc = launch())) return rc;
    }
    MPLB_CUDA(cudaMemcpyAsync(valid, dOut, n, cudaMemcpyDeviceToHost, c->stream));
    if (timeline && streamed) cudaEventRecord(tl[4], c->stream);
    int frc = finishCall(sc, c, n, nullptr);
    if (streamed) {
        cudaStreamSynchronize(c->copyStream);
        cudaStreamSynchronize(c->copyStream2);
    }
    if (timeline && streamed) {
        float b, k, d;
        cudaEventElapsedTime(&b, tl[0], tl[2]);
        cudaEventElapsedTime(&k, tl[1], tl[3]);
        cudaEventElapsedTime(&d, tl[1], tl[4]);
        std::fprintf(stderr, "[mplb timeline] n=%zu  copies done %.3f ms, kernel done %.3f, verdicts back %.3f, host %.3f\n",
                     n, b, k, d, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count());
    }
    if (streamed && c->hCtr->pad) {
        // The chunks did not reach the waiting kernel in time: something (a profiler replaying or serialising launches,
        // a debugger) keeps the copies from running under it.  This batch is repeated copy-then-kernel; two such calls
        // in a row switch the scene to that order until a later probe (every 64th large call) streams again.
        sc->hStreamStalls += 1;
        sc->streamStallsInARow.fetch_add(1, std::memory_order_relaxed);
        if (frc == MPLB_OK) sc->hChecks -= n;   // counted again below
        MPLB_CUDA(cudaMemsetAsync(c->ctr.ptr, 0, sizeof(DevCounters), c->stream));
        if ((rc = se3StatesPipelined(sc, c, states, n, dIn, dOut))) return rc;
        MPLB_CUDA(cudaMemcpyAsync(valid, dOut, n, cudaMemcpyDeviceToHost, c->stream));
        return finishCall(sc, c, n, nullptr);
    }
    if (streamed) {
        sc->streamStallsInARow.store(0, std::memory_order_relaxed);
    } else if (streamBatch(n) && (sc->unstreamedCalls.fetch_add(1, std::memory_order_relaxed) & 63u) == 63u) {
        sc->streamStallsInARow.store(1, std::memory_order_relaxed);   // probe: the next large call streams again
    }
    return frc;
}

static int se3CheckEdges(Se3Scene *sc, const double *from, const double *to, size_t n, uint8_t *valid,
                         uint64_t *statesChecked) {
    MPLB_CUDA(cudaSetDevice(sc->device));
    CallCtx *c = sc->acquire();
    if (!c) return fail(MPLB_ERR_CUDA, "could not create a call context (stream / pinned memory)");
    CtxGuard g{sc, c};
    c->hSteps.resize(n);
    uint32_t maxSteps = 0;
    static const bool timeline = std::getenv("MPLB_TIMELINE") != nullptr;
    const auto h0 = std::chrono::steady_clock::now();
    auto sinceMs = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count(); };
    // Large batches from page-locked memory: the end states start crossing the link before the step counts are planned
    // on the host (the copy engine needs no host thread), so the two overlap (65,536 edges: 0.11 + 0.16 ms -> 0.16 ms).
    bool copiedEarly = false;
    if (n > kSmallEdges && isPinned(from) && isPinned(to)) {
        const size_t bytes = n * 7 * sizeof(double);
        int rc0;
        if ((rc0 = c->in0.reserve(bytes)) || (rc0 = c->in1.reserve(bytes))) return rc0;
        if ((rc0 = hostToDevice(c, c->in0.ptr, from, bytes, c->copyStream)) || (rc0 = hostToDevice(c, c->in1.ptr, to, bytes, c->copyStream)))
            return rc0;
        if (c->events.empty()) {
            cudaEvent_t e;
            MPLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->events.push_back(e);
        }
        MPLB_CUDA(cudaEventRecord(c->events[0], c->copyStream));
        copiedEarly = true;
    }
    struct CopyQuiet {   // the copy stream is idle before the context goes back to the pool, whatever path the call leaves by
        CallCtx *c;
        bool on;
        ~CopyQuiet() {
            if (on) cudaStreamSynchronize(c->copyStream);
        }
    } copyQuiet{c, copiedEarly};
    {
        // se3...:139 on the host (same libm as the CPU path); large batches use the host's cores
        long long bad = -1;
        uint32_t *hs = c->hSteps.data();
#pragma omp parallel for schedule(static) if (n >= 8192) reduction(max : maxSteps)
        for (long long i = 0; i < (long long)n; ++i) {
            const uint64_t s = sc->edgeSteps(from + 7 * i, to + 7 * i);
            if (!(s < (1ull << 31))) {
#pragma omp critical
                bad = i;
                hs[i] = 0;
                continue;
            }
            hs[i] = (uint32_t)s;
            maxSteps = std::max(maxSteps, (uint32_t)s);
        }
        if (bad >= 0)
            return fail(MPLB_ERR_INVALID, "edge %lld needs >= 2^31 interpolation steps; non-finite state?", bad);
    }
    const double tPlan = sinceMs();
    int rc;
    const size_t sb = n * 7 * sizeof(double);
    const size_t stackBytes = edgeStackBytes(sc->numSms);
    if (n <= kSmallEdges) {
        if ((rc = c->aux0.reserve(stackBytes))) return rc;
        std::memcpy(c->hStage, from, sb);
        std::memcpy(c->hStage + sb, to, sb);
        std::memcpy(c->hStage + 2 * sb, c->hSteps.data(), n * 4);
        unsigned char *ctl = (unsigned char *)c->ctl.ptr, *in = (unsigned char *)c->stage.ptr;
        if (c->dStageHost) {
            // direct return (se3_device.cuh: HostReturn): the first kernel pulls the edges out of mapped host memory,
            // the last pass's last warp hands dead flags and counters back and rings; no copy engine, no stream sync
            if (c->ctlDirty) {
                MPLB_CUDA(cudaMemsetAsync(ctl, 0, kCtlHeader + kCtlTail, c->stream));
                c->ctlDirty = false;
            }
            HostReturn ret;
            ret.hostCtr = (DevCounters *)c->dCtlHost;
            ret.flag = (unsigned *)(c->dCtlHost + kCtlHeader - 64);
            // the call's number travels through pinned memory (the word after the flag), not through the launch
            // parameters: the scalar call's four kernels replay as one captured graph
            const unsigned seq = ++c->seq;
            *(volatile unsigned *)(c->hCtl + kCtlHeader - 60) = seq;
            ret.seqAt = (const unsigned *)(c->dCtlHost + kCtlHeader - 60);
            ret.clearWords = (unsigned)((kCtlHeader - sizeof(DevCounters)) / 4 + n);
            ret.tailAt = (unsigned)(kCtlHeader / 4);
            ret.tailWords = (unsigned)n;
            auto enqueueDirect = [&]() -> int {
                MPLB_CUDA(launchCheckEdges(sc->dev, (const double *)in, (const double *)(in + sb), (const unsigned *)(in + 2 * sb), n,
                                           (unsigned *)(ctl + kCtlHeader), nullptr, (DevCounters *)ctl,
                                           (unsigned long long *)(ctl + kCtlHeader - 128), c->aux0.ptr, stackBytes,
                                           (double *)(in + ((2 * sb + n * 4 + 7) & ~(size_t)7)), sc->numSms, c->stream, true, ret,
                                           c->dStageHost));
                return MPLB_OK;
            };
            if (n == 1) {
                const void *now[3] = {c->stage.ptr, c->ctl.ptr, c->aux0.ptr};
                if ((rc = runScalar(c, c->graphEdge, c->graphEdgeKey, now, 3, c->scalarEdges, enqueueDirect))) return rc;
            } else if ((rc = enqueueDirect())) {
                return rc;
            }
            sc->hLaunches += 4;
            if ((rc = awaitRing(c, seq))) return rc;
            const unsigned *dead = (const unsigned *)(c->hCtl + kCtlHeader);
            for (size_t i = 0; i < n; ++i) valid[i] = dead[i] ? 0 : 1;
            rc = account(sc, *(const DevCounters *)c->hCtl, 0, statesChecked);
            if (timeline)
                std::fprintf(stderr, "[mplb timeline] %zu edges (small call, direct return): steps planned %.3f ms, done %.3f\n", n, tPlan, sinceMs());
            return rc;
        }
        c->ctlDirty = true;
        auto enqueue = [&]() -> int {
            MPLB_CUDA(cudaMemsetAsync(ctl, 0, kCtlHeader + n * 4, c->stream));
            MPLB_CUDA(cudaMemcpyAsync(in, c->hStage, 2 * sb + n * 4, cudaMemcpyHostToDevice, c->stream));
            MPLB_CUDA(launchCheckEdges(sc->dev, (const double *)in, (const double *)(in + sb), (const unsigned *)(in + 2 * sb), n,
                                       (unsigned *)(ctl + kCtlHeader), nullptr, (DevCounters *)ctl,
                                       (unsigned long long *)(ctl + kCtlHeader - 128), c->aux0.ptr, stackBytes,
                                       (double *)(in + ((2 * sb + n * 4 + 7) & ~(size_t)7)), sc->numSms, c->stream, true));
            MPLB_CUDA(cudaMemcpyAsync(c->hCtl, ctl, kCtlHeader + n * 4, cudaMemcpyDeviceToHost, c->stream));
            return MPLB_OK;
        };
        if (n == 1) {
            const void *now[3] = {c->stage.ptr, c->ctl.ptr, c->aux0.ptr};
            if ((rc = runScalar(c, c->graphEdge, c->graphEdgeKey, now, 3, c->scalarEdges, enqueue))) return rc;
        } else if ((rc = enqueue())) {
            return rc;
        }
        sc->hLaunches += 4;   // per-edge constants + three passes
        MPLB_CUDA(cudaStreamSynchronize(c->stream));
        const unsigned *dead = (const unsigned *)(c->hCtl + kCtlHeader);
        for (size_t i = 0; i < n; ++i) valid[i] = dead[i] ? 0 : 1;
        rc = account(sc, *(const DevCounters *)c->hCtl, 0, statesChecked);
        if (timeline)
            std::fprintf(stderr, "[mplb timeline] %zu edges (small call): steps planned %.3f ms, done %.3f\n", n, tPlan, sinceMs());
        return rc;
    }
    if ((rc = c->in0.reserve(sb)) || (rc = c->in1.reserve(sb)) || (rc = c->steps.reserve(n * 4)) ||
        (rc = c->dead.reserve(n * 4)) || (rc = c->out.reserve(n)) || (rc = c->aux0.reserve(stackBytes)) ||
        (rc = c->aux1.reserve(edgeConstBytes(n))))
        return rc;
    static thread_local cudaEvent_t tl[4] = {};
    if (timeline && !tl[0])
        for (auto &e : tl) cudaEventCreate(&e);
    if (timeline) cudaEventRecord(tl[0], c->stream);
    MPLB_CUDA(cudaMemsetAsync(c->ctr.ptr, 0, sizeof(DevCounters), c->stream));
    if (copiedEarly) {
        MPLB_CUDA(cudaStreamWaitEvent(c->stream, c->events[0], 0));
    } else if ((rc = hostToDevice(c, c->in0.ptr, from, sb, c->stream)) || (rc = hostToDevice(c, c->in1.ptr, to, sb, c->stream))) {
        return rc;
    }
    MPLB_CUDA(cudaMemcpyAsync(c->steps.ptr, c->hSteps.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
    if (timeline) cudaEventRecord(tl[1], c->stream);
    MPLB_CUDA(launchCheckEdges(sc->dev, (const double *)c->in0.ptr, (const double *)c->in1.ptr,
                               (const unsigned *)c->steps.ptr, n, (unsigned *)c->dead.ptr, (uint8_t *)c->out.ptr,
                               (DevCounters *)c->ctr.ptr, (unsigned long long *)c->work.ptr, c->aux0.ptr, stackBytes,
                               (double *)c->aux1.ptr, sc->numSms, c->stream));
    sc->hLaunches += 5;   // per-edge constants, three passes, the verdict kernel
    if (timeline) cudaEventRecord(tl[2], c->stream);
    const double tLaunched = sinceMs();
    MPLB_CUDA(cudaMemcpyAsync(valid, c->out.ptr, n, cudaMemcpyDeviceToHost, c->stream));
    if (timeline) cudaEventRecord(tl[3], c->stream);
    const double tEnq = sinceMs();
    rc = finishCall(sc, c, 0, statesChecked);
    if (timeline) {
        float a, b, d;
        cudaEventElapsedTime(&a, tl[0], tl[1]);
        cudaEventElapsedTime(&b, tl[1], tl[2]);
        cudaEventElapsedTime(&d, tl[2], tl[3]);
        std::fprintf(stderr, "[mplb timeline] %zu edges: steps planned %.3f ms, launched %.3f, enqueued %.3f, done %.3f; device: copies in %.3f, kernels %.3f, verdicts out %.3f\n",
                     n, tPlan, tLaunched, tEnq, sinceMs(), a, b, d);
    }
    return rc;
}

int se3CheckEdgesHost(Se3Scene *sc, const double *from, const double *to, size_t n, uint8_t *valid, uint64_t *statesChecked) {
    return se3CheckEdges(sc, from, to, n, valid, statesChecked);
}

int Se3Scene::checkStates(const double *states, size_t n, uint8_t *valid) { return se3CheckStates(this, states, n, valid); }
int Se3Scene::checkEdges(const double *from, const double *to, size_t n, uint8_t *valid, uint64_t *statesChecked) {
    return se3CheckEdges(this, from, to, n, valid, statesChecked);
}

}  // namespace mplb

using namespace mplb;

static SceneBase *base(mplb_scene *s) { return reinterpret_cast<SceneBase *>(s); }
static const SceneBase *base(const mplb_scene *s) { return reinterpret_cast<const SceneBase *>(s); }

extern "C" {

const char *mplb_last_error(void) { return gLastError.c_str(); }

int mplb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void mplb_free(void *p) { std::free(p); }

int mplb_set_cache_dir(const char *dir) {
    if (setCacheDir(dir) != 0) return fail(MPLB_ERR_IO, "cannot use %s as the cache directory", dir ? dir : "(null)");
    return MPLB_OK;
}

void mplb_cache_stats(uint64_t *hits, uint64_t *misses) { cacheStats(hits, misses); }

int mplb_mesh_load_collada(const char *path, int shift_to_center, int identity_root_transform, double **tris,
                           size_t *ntris, double center[3]) {
    if (!path || !tris || !ntris) return fail(MPLB_ERR_INVALID, "null argument");
    *tris = nullptr;
    *ntris = 0;
    try {
        TriangleSoup s = loadColladaCached(path, shift_to_center != 0, identity_root_transform != 0);
        double *buf = (double *)std::malloc(std::max<size_t>(1, s.tris.size()) * sizeof(double));
        if (!buf) return fail(MPLB_ERR_NOMEM, "out of memory");
        std::memcpy(buf, s.tris.data(), s.tris.size() * sizeof(double));
        *tris = buf;
        *ntris = s.size();
        if (center) std::memcpy(center, s.center, sizeof(s.center));
        return MPLB_OK;
    } catch (const std::invalid_argument &e) {
        return fail(MPLB_ERR_IO, "%s", e.what());
    } catch (const std::exception &e) {
        return fail(MPLB_ERR_IO, "%s", e.what());
    }
}

int mplb_scene_create_se3(const double *env_tris, size_t n_env, const double *robot_tris, size_t n_robot,
                          double so3_weight, double l2_weight, double check_resolution, int device,
                          mplb_scene **out) {
    try {
        return createSe3(env_tris, n_env, robot_tris, n_robot, so3_weight, l2_weight, check_resolution, device, out);
    } catch (const std::exception &e) {
        return fail(MPLB_ERR_INVALID, "%s", e.what());
    }
}

int mplb_scene_create_se3_files(const char *env_path, const char *robot_path, double so3_weight, double l2_weight,
                                double check_resolution, int device, mplb_scene **out) {
    if (!env_path || !robot_path) return fail(MPLB_ERR_INVALID, "null path");
    try {
        // se3_rigid_body_scenario.hpp:58-59: env load(false,false), robot load(true,false)
        TriangleSoup env = loadColladaCached(env_path, false, false);
        TriangleSoup robot = loadColladaCached(robot_path, true, false);
        return createSe3(env.tris.data(), env.size(), robot.tris.data(), robot.size(), so3_weight, l2_weight,
                         check_resolution, device, out);
    } catch (const std::invalid_argument &e) {
        return fail(MPLB_ERR_IO, "%s", e.what());
    } catch (const std::exception &e) {
        return fail(MPLB_ERR_INVALID, "%s", e.what());
    }
}

void mplb_scene_destroy(mplb_scene *scene) {
    if (!scene) return;
    SceneBase *b = base(scene);
    cudaSetDevice(b->device);
    cudaDeviceSynchronize();
    delete b;
}

static bool coalescing(const SceneBase *b) {
    static const bool envOn = std::getenv("MPLB_COALESCE") != nullptr;
    const int c = b->coalesce.load(std::memory_order_relaxed);
    return c < 0 ? envOn : c != 0;
}

int mplb_scene_state_dim(const mplb_scene *scene) { return scene ? base(scene)->stateDim : 0; }

int mplb_check_states(mplb_scene *scene, const double *states, size_t n, uint8_t *valid) {
    if (!scene) return fail(MPLB_ERR_INVALID, "scene is null");
    if (n == 0) return MPLB_OK;
    if (!states || !valid) return fail(MPLB_ERR_INVALID, "null buffer");
    if (n == 1 && coalescing(base(scene))) return base(scene)->combined(0, states, nullptr, valid);
    return base(scene)->checkStates(states, n, valid);
}

int mplb_check_edges(mplb_scene *scene, const double *from, const double *to, size_t n, uint8_t *valid,
                     uint64_t *states_checked) {
    if (!scene) return fail(MPLB_ERR_INVALID, "scene is null");
    if (states_checked) *states_checked = 0;
    if (n == 0) return MPLB_OK;
    if (!from || !to || !valid) return fail(MPLB_ERR_INVALID, "null buffer");
    if (n == 1 && !states_checked && coalescing(base(scene)))
        return base(scene)->combined(1, from, to, valid);
    return base(scene)->checkEdges(from, to, n, valid, states_checked);
}

int mplb_check_states_device(mplb_scene *scene, const double *d_states, size_t n, uint8_t *d_valid, void *stream) {
    if (!scene) return fail(MPLB_ERR_INVALID, "scene is null");
    if (n == 0) return MPLB_OK;
    if (!d_states || !d_valid) return fail(MPLB_ERR_INVALID, "null buffer");
    auto *s = dynamic_cast<Se3Scene *>(base(scene));
    if (!s) return fail(MPLB_ERR_INVALID, "unsupported scene kind");
    MPLB_CUDA(cudaSetDevice(s->device));
    {
        std::lock_guard<std::mutex> g(s->mu);
        const cudaStream_t st = (cudaStream_t)stream;
        if (!s->cursorWords.ptr) {
            int rc = s->cursorWords.reserve((Se3Scene::kDeviceCursors + 1) * 64);
            if (rc != MPLB_OK) return rc;
        }
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        MPLB_CUDA(cudaStreamIsCapturing(st, &cap));
        unsigned slot = Se3Scene::kDeviceCursors;   // a captured call: ordered by its graph, no events across the capture
        if (cap == cudaStreamCaptureStatusNone) {
            slot = s->nextCursor++ % Se3Scene::kDeviceCursors;
            Se3Scene::DeviceCursor &c = s->cursors[slot];
            if (!c.done) MPLB_CUDA(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
            if (c.used) MPLB_CUDA(cudaStreamWaitEvent(st, c.done, 0));
        }
        MPLB_CUDA(launchCheckStates(s->dev, d_states, n, d_valid, (DevCounters *)s->globalCtr.ptr,
                                    (unsigned *)((char *)s->cursorWords.ptr + 64 * slot), s->numSms, st));
        if (slot < Se3Scene::kDeviceCursors) {
            MPLB_CUDA(cudaEventRecord(s->cursors[slot].done, st));
            s->cursors[slot].used = true;
        }
    }
    s->hLaunches += 1;
    s->hChecks += n;
    return MPLB_OK;
}

int mplb_edges_plan(mplb_scene *scene, const double *from, const double *to, size_t n, uint32_t *steps) {
    if (!scene) return fail(MPLB_ERR_INVALID, "scene is null");
    if (n == 0) return MPLB_OK;
    if (!from || !to || !steps) return fail(MPLB_ERR_INVALID, "null buffer");
    return base(scene)->planEdges(from, to, n, steps);
}

int mplb_check_edges_device(mplb_scene *scene, const double *d_from, const double *d_to, const uint32_t *d_steps,
                            size_t n, uint32_t max_steps, uint8_t *d_valid, void *stream) {
    if (!scene) return fail(MPLB_ERR_INVALID, "scene is null");
    if (n == 0) return MPLB_OK;
    if (!d_from || !d_to || !d_steps || !d_valid) return fail(MPLB_ERR_INVALID, "null buffer");
    auto *s = dynamic_cast<Se3Scene *>(base(scene));
    if (!s) return fail(MPLB_ERR_INVALID, "unsupported scene kind");
    MPLB_CUDA(cudaSetDevice(s->device));
    {
        // the dead-flag scratch is shared by device-side callers: serialise their use of it
        std::lock_guard<std::mutex> g(s->mu);
        int rc = s->globalDead.reserve(n * 4);
        if (rc != MPLB_OK) return rc;
        if ((rc = s->globalStacks.reserve(edgeStackBytes(s->numSms))) != MPLB_OK) return rc;
        if ((rc = s->globalEdgeConst.reserve(edgeConstBytes(n))) != MPLB_OK) return rc;
    }
    (void)max_steps;
    MPLB_CUDA(launchCheckEdges(s->dev, d_from, d_to, d_steps, n, (unsigned *)s->globalDead.ptr, d_valid,
                               (DevCounters *)s->globalCtr.ptr, (unsigned long long *)s->globalWork.ptr,
                               s->globalStacks.ptr, edgeStackBytes(s->numSms), (double *)s->globalEdgeConst.ptr, s->numSms,
                               (cudaStream_t)stream));
    s->hLaunches += 5;
    return MPLB_OK;
}

double mplb_distance(const mplb_scene *scene, const double *a, const double *b) {
    if (!scene || !a || !b) return NAN;
    return base(scene)->distance(a, b);
}

int mplb_scene_track_margins(mplb_scene *scene, int on) {
    if (!scene) return fail(MPLB_ERR_INVALID, "scene is null");
    auto *s = dynamic_cast<Se3Scene *>(base(scene));
    if (!s) return fail(MPLB_ERR_INVALID, "margins are tracked for SE(3) scenes");
    MPLB_CUDA(cudaSetDevice(s->device));
    MPLB_CUDA(cudaDeviceSynchronize());
    if (on && !s->marginWord.ptr) {
        int rc = s->marginWord.reserve(sizeof(unsigned));
        if (rc) return rc;
        MPLB_CUDA(cudaMemset(s->marginWord.ptr, 0, sizeof(unsigned)));
    }
    s->dev.minMargin = on ? (unsigned *)s->marginWord.ptr : nullptr;
    return MPLB_OK;
}

int mplb_scene_set_coalescing(mplb_scene *scene, int on) {
    if (!scene) return fail(MPLB_ERR_INVALID, "scene is null");
    base(scene)->coalesce.store(on ? 1 : 0, std::memory_order_relaxed);
    return MPLB_OK;
}

int mplb_scene_stats(mplb_scene *scene, mplb_stats_t *out, int reset) {
    if (!scene || !out) return fail(MPLB_ERR_INVALID, "null argument");
    SceneBase *b = base(scene);
    MPLB_CUDA(cudaSetDevice(b->device));
    MPLB_CUDA(cudaDeviceSynchronize());
    DevCounters g{};
    MPLB_CUDA(cudaMemcpy(&g, b->globalCtr.ptr, sizeof(g), cudaMemcpyDeviceToHost));
    if (g.overflow) return fail(MPLB_ERR_CUDA, "device traversal bookkeeping error (pair stack / slot accounting)");
    std::memset(out, 0, sizeof(*out));
    out->checks = b->hChecks + g.states;
    out->bv_tests = b->hBv + g.bvTests;
    out->tri_tests = b->hTri + g.triTests;
    out->exact_tests = b->hExact + g.exactTests;
    out->launches = b->hLaunches;
    out->env_nodes = b->envNodeCount;
    out->robot_nodes = b->robotNodeCount;
    out->stream_stalls = b->hStreamStalls;
    out->ambiguous_steps = b->hAmbiguous;
    out->coalesced_calls = b->hCoalesced;
    out->min_exact_margin = INFINITY;
    if (auto *s = dynamic_cast<Se3Scene *>(b)) {
        if (s->marginWord.ptr) {
            unsigned w = 0;
            MPLB_CUDA(cudaMemcpy(&w, s->marginWord.ptr, sizeof(w), cudaMemcpyDeviceToHost));
            if (w) {
                const unsigned bits = ~w;
                float f;
                std::memcpy(&f, &bits, sizeof(f));
                out->min_exact_margin = f;
            }
            if (reset) MPLB_CUDA(cudaMemset(s->marginWord.ptr, 0, sizeof(unsigned)));
        }
        out->env_tris = s->nEnvTris;
        out->robot_tris = s->nRobotTris;
        out->streaming_off = s->streamStallsInARow.load(std::memory_order_relaxed) >= 2;
    }
    if (reset) {
        MPLB_CUDA(cudaMemset(b->globalCtr.ptr, 0, sizeof(DevCounters)));
        b->hChecks = b->hBv = b->hTri = b->hExact = b->hLaunches = 0;
        b->hStreamStalls = b->hAmbiguous = b->hCoalesced = 0;
    }
    return MPLB_OK;
}

}  // extern "C"
