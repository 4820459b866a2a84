// Internal host-side state behind the opaque mplb_scene handle.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <memory>
#include <cmath>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/mplb.h"
#include "se3_device.cuh"

namespace mplb {

extern thread_local std::string gLastError;
int fail(int code, const char *fmt, ...);

struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);  // grows only; contents are not preserved
    ~DevBuf();
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

// Everything one in-flight host call needs; borrowed from the scene so that concurrent callers
// (the reference's OpenMP planner threads, prrt.hpp:140-152) never share a stream or a buffer.
struct CallCtx {
    cudaStream_t stream = nullptr;      // kernels + verdict read-back
    cudaStream_t copyStream = nullptr;  // host->device input copies of pipelined batches
    cudaStream_t copyStream2 = nullptr; // streamed batches alternate their chunks over both copy streams
    std::vector<cudaEvent_t> events;    // one per pipeline stage
    DevBuf in0, in1, steps, dead, out, ctr, work, aux0, aux1, aux2, aux3, aux4, aux5, aux6;
    DevBuf ivA, ivB, ivMisc;   // Fetch edge batches: the two interval lists, and motion / forced flags / counts
    DevCounters *hCtr = nullptr;  // pinned
    // Small calls (the reference's scalar isValid) go through one pinned staging buffer each way and one
    // device control block [DevCounters | 16 work counters | verdicts or dead flags]: one memset, one
    // host->device copy, the kernels, one device->host copy.
    unsigned char *hStage = nullptr, *hCtl = nullptr;   // pinned
    DevBuf stage, ctl;
    // Direct return (se3_device.cuh: HostReturn): the device's view of the two pinned buffers (null: not mapped), the
    // call counter the kernel rings back, and whether the device control block still has to be cleared (it is left
    // clean by every direct-return call; the older copy-in / copy-out sequences clear it themselves on entry)
    unsigned char *dStageHost = nullptr, *dCtlHost = nullptr;
    unsigned seq = 0;
    bool ctlDirty = true;
    // large inputs in pageable memory: two pinned bounce buffers filled by all cores while the previous one is on
    // the wire (cudaMemcpyAsync on pageable memory stages through one thread at ~10 GB/s)
    unsigned char *hBounce[2] = {nullptr, nullptr};
    cudaEvent_t bounceDone[2] = {nullptr, nullptr};
    int bounceNext = 0;                 // the bounce buffer the next piece goes through
    // streamed batches (capi.cu): one arrival flag per chunk, written behind the chunk's copy with a stream memory
    // operation; `epoch` names the call, so the flags never need clearing
    DevBuf flags;
    unsigned epoch = 0;
    // The reference's scalar calls (one state, one edge) replay a captured CUDA graph: memset + copy in + kernels +
    // copy out are one launch instead of four to eight.  Captured on a context's second scalar call; the keys are the
    // device buffers the graph was captured with (a reallocation invalidates it).
    cudaGraphExec_t graphState = nullptr, graphEdge = nullptr;
    void *graphStateKey[2] = {nullptr, nullptr}, *graphEdgeKey[3] = {nullptr, nullptr, nullptr};
    unsigned scalarStates = 0, scalarEdges = 0;
    bool graphsOff = false;
    // pinned result buffers of pipelined rounds (Fetch edge intervals), grown on demand
    unsigned char *hRes[2] = {nullptr, nullptr};
    size_t hResCap[2] = {0, 0};
    int reserveRes(int i, size_t bytes);
    unsigned long long *hWater = nullptr;  // pinned: per-slice arrival marks of a streamed batch
    std::vector<uint32_t> hSteps;
    ~CallCtx();
};

// Host -> device copy of a caller's buffer on `stream`: pinned memory goes straight to the copy engine, pageable memory
// through two pinned bounce buffers filled by all cores (capi.cu)
int hostToDevice(CallCtx *c, void *dst, const void *src, size_t bytes, cudaStream_t stream);
// true when `p` is page-locked host memory the copy engine can read directly
bool isPinned(const void *p);

struct SceneBase {
    int device = 0;
    int numSms = 0;
    int stateDim = 0;
    std::mutex mu;
    std::vector<CallCtx *> freeCtx;
    DevBuf globalCtr, globalWork, globalDead, globalStacks, globalEdgeConst;  // used by the *_device entry points
    std::atomic<uint64_t> hChecks{0}, hBv{0}, hTri{0}, hExact{0}, hLaunches{0};
    std::atomic<uint64_t> hStreamStalls{0}, hAmbiguous{0}, hCoalesced{0};
    size_t envNodeCount = 0, robotNodeCount = 0;

    // Scalar calls -- the reference's isValid(q) / isValid(from, to), one per OpenMP planner thread (prrt.hpp:140-152,
    // 280-299) -- can share launches (capi.cu: combined()): calls that arrive while earlier ones are on the device ride
    // together, a lone caller is never delayed.  Off by default: every caller already has its own context and stream,
    // the device runs their one-CTA kernels side by side, and with T threads at most T calls are in flight, so sharing
    // a launch only adds the wait for it (measured, 16 threads: 408 k calls/s apart, 274 k shared).  It pays when
    // launches, not latency, are what is short: more threads than the driver launches concurrently.  [0] states, [1] edges.
    struct Combiner {
        static constexpr int kMax = 64;   // calls per launch
        struct Batch {
            std::vector<double> a, b;     // [kMax][stateDim] states (or edge starts), edge ends
            uint8_t out[kMax];
            int n = 0;                    // under Combiner::mu
            std::atomic<int> filled{0};   // items whose payload has been copied in
            std::atomic<int> left{0};     // callers that have read their verdict
            std::atomic<int> done{0};     // 1: verdicts are there, 2: the launch failed (every caller repeats its own call)
        };
        std::mutex mu;
        std::shared_ptr<Batch> open;      // the batch newcomers join (null: the next caller starts one)
        std::vector<std::shared_ptr<Batch>> spare;
        std::atomic<int> inflight{0};     // batches on the device
    } comb[2];
    std::atomic<int> coalesce{-1};    // -1: default (off unless MPLB_COALESCE is set), 0 / 1: mplb_scene_set_coalescing
    int combined(int kind, const double *a, const double *b, uint8_t *valid);

    CallCtx *acquire();
    void release(CallCtx *c);
    virtual ~SceneBase();
    // per-kind implementations of the generic entry points (capi.cu dispatches through these)
    virtual int checkStates(const double *states, size_t n, uint8_t *valid) = 0;
    virtual int checkEdges(const double *from, const double *to, size_t n, uint8_t *valid, uint64_t *statesChecked) = 0;
    virtual double distance(const double *a, const double *b) const = 0;
    virtual int planEdges(const double *from, const double *to, size_t n, uint32_t *steps) const = 0;
};

// the SE(3) rigid-body scene (capi.cu; the device-resident planner loop of planner_capi.cu works on it too)
struct Se3Scene : SceneBase {
    double so3w = 50, l2w = 1, invStep = 10;
    // streamed batches whose input did not arrive under the kernel, in a row: two switch streaming off (capi.cu)
    std::atomic<unsigned> streamStallsInARow{0}, unstreamedCalls{0};
    SceneDev dev{};
    DevBuf robotNodes, envNodes, robotTris, envTris, robotTris64, envTris64;
    DevBuf marginWord;   // mplb_scene_track_margins: the device word SceneDev::minMargin points at
    size_t nEnvTris = 0, nRobotTris = 0;
    // mplb_check_states_device: every call in flight needs a work cursor of its own -- callers on different streams run
    // side by side, and the SMs a launch has drained take the next launch's CTAs while its last warps finish.  A ring of
    // cursors; a cursor is handed out again once the call that used it four calls ago is through (an event the new
    // call's stream waits for -- a no-op for a caller that stays on one stream).
    struct DeviceCursor {
        cudaEvent_t done = nullptr;
        bool used = false;
    };
    static constexpr unsigned kDeviceCursors = 4;
    DeviceCursor cursors[kDeviceCursors];
    DevBuf cursorWords;   // kDeviceCursors + 1 cursors, 64 bytes apart (the last one: calls on a capturing stream)
    unsigned nextCursor = 0;
    ~Se3Scene() override {
        for (auto &c : cursors)
            if (c.done) cudaEventDestroy(c.done);
    }

    Se3Scene() { stateDim = 7; }
    int checkStates(const double *states, size_t n, uint8_t *valid) override;
    int checkEdges(const double *from, const double *to, size_t n, uint8_t *valid, uint64_t *statesChecked) override;
    int planEdges(const double *from, const double *to, size_t n, uint32_t *steps) const override {
        for (size_t i = 0; i < n; ++i) {
            uint64_t st = edgeSteps(from + 7 * i, to + 7 * i);
            if (!(st < (1ull << 31))) return fail(MPLB_ERR_INVALID, "edge %zu: step count out of range", i);
            steps[i] = (uint32_t)st;
        }
        return MPLB_OK;
    }

    double distance(const double *a, const double *b) const override {
        // nigh SE3Space<S,so3w,l2w> [EXT]: so3w * acos(|qa . qb|) + l2w * ||ta - tb||
        double d = std::fabs(((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) + a[3] * b[3]);
        double so3 = d < 1.0 ? std::acos(d) : 0.0;
        double dx = a[4] - b[4], dy = a[5] - b[5], dz = a[6] - b[6];
        return so3w * so3 + l2w * std::sqrt((dx * dx + dy * dy) + dz * dz);
    }
    // se3_rigid_body_scenario.hpp:139
    uint64_t edgeSteps(const double *a, const double *b) const {
        // (a non-finite or huge distance must not reach the integer conversion: that is undefined, and yields 0 for
        // +inf with the usual x86 sequence -- the callers reject anything >= 2^31)
        const double x = std::ceil(distance(a, b) * invStep);
        return x < 9.0e18 ? (uint64_t)x : ~0ull;
    }
};

struct CtxGuard {
    SceneBase *sc;
    CallCtx *c;
    ~CtxGuard() {
        if (c) sc->release(c);
    }
};

#define MPLB_CUDA(expr)                                                                             \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return ::mplb::fail(-3, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int selectDevice(int device, int *numSms);
int bitsFor(size_t n);

template <typename T>
int upload(DevBuf &buf, const std::vector<T> &v) {
    int rc = buf.reserve(v.size() * sizeof(T) + 16);
    if (rc != 0) return rc;
    if (!v.empty()) MPLB_CUDA(cudaMemcpy(buf.ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}


}  // namespace mplb
