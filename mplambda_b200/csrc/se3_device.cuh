// Device-side scene description and launch entry points shared by the C-ABI and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace mplb {

// Flat, immutable, HBM-resident scene (built once on the host by bvh_build.cpp).
struct SceneDev {
    const float4 *robotNodes;   // 4 x float4 per node (mplb_host.hpp: BvhNode)
    const float4 *envNodes;
    const float4 *robotTris;    // 4 x float4 per triangle (FP32 filter copy; the 4th is padding)
    const float4 *envTris;
    const double *robotTris64;  // 9 doubles per triangle (exact re-evaluation)
    const double *envTris64;
    int nRobotNodes, nEnvNodes;
    int haveGeometry;           // both meshes non-empty
    int bothRootsLeaf;          // both meshes are a single triangle
    int bitsA, bitsB;           // bits needed for a robot / env node index
    int wide;                   // 1: pair entries need 64 bits
    float envRadius;            // max |vertex| of the environment
    float robotRadius;          // max |vertex| of the (centred) robot
    int triBatchMin;            // leaf pairs queued before a triangle batch runs (<= 32)
    int triEagerLanes;          // ... or any queued pair once at most this many lanes still search
    int stackLevels;            // per-lane LIFO levels = depthA + depthB + 2 (rounded up)
    unsigned rootEntry;         // LIFO entry that expands the root pair (se3_kernels.cu)
    unsigned *minMargin;        // diagnostic (mplb_scene_track_margins): ~bits of the smallest relative margin of an exact
                                // triangle-pair decision so far; null: not tracked (the exact path exits early)
};

constexpr unsigned kAmbiguousListed = 14;   // ambiguous edges listed by index (more are only counted)

struct DevCounters {
    unsigned long long bvTests, triTests, exactTests, states;
    unsigned overflow, pad;   // pad: 1 = a streamed batch's input never arrived (see StateSource)
    // device-planned edges (edgeGatherPlanKernel): distance * invStep within a few ulp of an integer, where CUDA's and
    // glibc's acos could round `steps` differently -- the caller re-plans those with host libm; badEdges: non-finite
    unsigned ambiguous, badEdges;
    unsigned ambiguousEdge[kAmbiguousListed];
    unsigned ticket, ticketPad;   // warps that have finished (direct-return calls, see HostReturn)
#ifdef MPLB_WARP_TIMES   // diagnostic build: when did the warps start / run out of new work / finish (globaltimer, ns)
    unsigned long long negStart, negFirstDry, negFirstDone, lastDone;   // neg*: ~t, so that atomicMax finds the minimum
    unsigned long long warpDone[4096], warpDry[4096];
    unsigned warpStates[4096], warpBv[4096], warpTri[4096];
#endif
};

// A batch that is still being copied in while the kernel runs: the host copies it in pieces on one stream and writes
// *water = epoch << kWaterShift | (states landed so far) >> kWaterUnitBits behind each piece (a stream memory
// operation); the kernel hands states out in index order and picks each claim up once the watermark covers it
// (se3_kernels.cu: StateSource).  Pieces end on multiples of 1024 states (the last one is rounded up), and the epoch
// changes with every call on the word, so the word never needs clearing.  water == nullptr: the batch is resident.
constexpr unsigned kWaterShift = 22, kWaterUnitBits = 10;
// Small host calls (the reference's scalar isValid and small planner batches) skip the copy engine both ways: the kernel
// reads the states from mapped page-locked host memory, writes the verdicts to mapped host memory, and its last warp
// copies the counters to the host, clears the device control block for the next call (counters, work counters, ticket)
// and rings the host: *flag = seq, a release store at system scope the calling thread spins on.  One kernel launch per
// call instead of memset + copy in + kernel + copy out + stream synchronisation.
struct HostReturn {
    DevCounters *hostCtr = nullptr;      // mapped host memory
    unsigned *flag = nullptr;            // mapped host memory
    unsigned seq = 0;
    const unsigned *seqAt = nullptr;     // non-null: the value to ring with is read from here (mapped host memory) when
                                         // the kernel ends -- the launch parameters stay the same from call to call, so
                                         // a captured graph of the call can be replayed
    unsigned clearWords = 0;             // 32-bit words to clear after the DevCounters block (work counters, dead flags)
    unsigned tailAt = 0, tailWords = 0;  // words [tailAt, tailAt + tailWords) of the control block (an edge call's dead
                                         // flags) are copied to the same place behind hostCtr before they are cleared
};

struct StreamedInput {
    const double *host = nullptr;   // zero-copy: the batch in mapped page-locked host memory (device-visible address);
                                    // the kernel pulls claims over PCIe itself and mirrors them into dStates
    const unsigned *water = nullptr;
    unsigned epoch = 0;
    HostReturn ret;                 // flag != nullptr: a direct-return call
};
// workZeroed: the caller has already zeroed the first work counter
cudaError_t launchCheckStates(const SceneDev &sc, const double *dStates, size_t n, uint8_t *dValid,
                              DevCounters *dCtr, unsigned *dWork, int numSms, cudaStream_t stream,
                              StreamedInput streamed = StreamedInput(), bool workZeroed = false);

// scratch for the per-warp interval LIFOs of checkEdgesKernel
size_t edgeStackBytes(int numSms);

cudaError_t launchCheckEdges(const SceneDev &sc, const double *dFrom, const double *dTo, const unsigned *dSteps,
                             size_t n, unsigned *dDead, uint8_t *dValid, DevCounters *dCtr, unsigned long long *dWork,
                             void *dStacks, size_t stackBytes, double *dEdgeConst, int numSms, cudaStream_t stream,
                             bool zeroed = false, HostReturn ret = HostReturn(), const void *hostInputs = nullptr);
// hostInputs (direct-return calls): from [n][7] | to [n][7] | steps [n] in mapped host memory; the first kernel copies
// them to dFrom / dTo / dSteps (device memory the passes read again and again) while it forms the per-edge constants
// dEdgeConst: edgeConstBytes(n) of scratch for the per-edge interpolation constants
size_t edgeConstBytes(size_t n);

// Edges named by index into state pools (device memory): edge e runs from row (fromIdx ? fromIdx[e] : e / fromDiv)
// of fromPool to row (toIdx ? toIdx[e] : e / toDiv) of toPool.
struct EdgeGather {
    const double *fromPool, *toPool;       // [rows][7]
    const long long *fromIdx, *toIdx;      // [n] or null
    unsigned long long fromRows, toRows;   // pool sizes: an index outside [0, rows) makes the edge invalid
    unsigned fromDiv, toDiv;               // >= 1
    unsigned forceAmbiguousEvery = 0;      // test hook: report every k-th edge as ambiguous (0: off)
};
// scratch the indexed call needs besides the interval stacks: dense end states, steps, per-edge constants
size_t edgeGatherBytes(size_t n);
// dScratch: edgeGatherBytes(n).  Verdicts as launchCheckEdges.  Steps are planned on the device (se3_kernels.cu:
// edgeGatherPlanKernel); ambiguous ones are reported in dCtr.
cudaError_t launchCheckEdgesIndexed(const SceneDev &sc, const EdgeGather &g, size_t n, double so3w, double l2w,
                                    double invStep, unsigned *dDead, uint8_t *dValid, DevCounters *dCtr,
                                    unsigned long long *dWork, void *dStacks, size_t stackBytes, void *dScratch,
                                    int numSms, cudaStream_t stream, bool zeroed = false);
// where the indexed call keeps edge e's planned step count / dense end states inside dScratch (for host re-planning)
unsigned *edgeGatherSteps(void *dScratch, size_t n);
double *edgeGatherFrom(void *dScratch, size_t n);
double *edgeGatherTo(void *dScratch, size_t n);
// zeroed: dWork (16 counters) and dDead are already zero.  dValid == nullptr: leave the verdicts as dDead flags.

}  // namespace mplb
