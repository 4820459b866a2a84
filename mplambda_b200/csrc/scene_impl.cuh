// Internal host-side state behind the opaque mplb_scene handle.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>
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
    std::vector<cudaEvent_t> events;    // one per pipeline stage
    DevBuf in0, in1, steps, dead, out, ctr, work, aux0, aux1;
    DevCounters *hCtr = nullptr;  // pinned
    std::vector<uint32_t> hSteps;
    ~CallCtx();
};

struct SceneBase {
    int device = 0;
    int numSms = 0;
    int stateDim = 0;
    std::mutex mu;
    std::vector<CallCtx *> freeCtx;
    DevBuf globalCtr, globalWork, globalDead;  // used by the *_device entry points
    std::atomic<uint64_t> hChecks{0}, hBv{0}, hTri{0}, hExact{0}, hLaunches{0};
    size_t envNodeCount = 0, robotNodeCount = 0;

    CallCtx *acquire();
    void release(CallCtx *c);
    virtual ~SceneBase();
};

}  // namespace mplb
