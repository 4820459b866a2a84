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
};

struct DevCounters {
    unsigned long long bvTests, triTests, exactTests, states;
    unsigned overflow, pad;   // pad: 1 = a streamed batch's input never arrived (see StateSource)
#ifdef MPLB_WARP_TIMES   // diagnostic build: when did the warps start / run out of new work / finish (globaltimer, ns)
    unsigned long long negStart, negFirstDry, negFirstDone, lastDone;   // neg*: ~t, so that atomicMax finds the minimum
    unsigned long long warpDone[4096];
    unsigned warpStates[4096], warpBv[4096], warpTri[4096];
#endif
};

// streamed: dStates was pre-filled with the all-ones pattern and is being overwritten by a host->device copy
// on another stream; the kernel picks each state up when it has landed (se3_kernels.cu: StateSource)
// workZeroed: the caller has already zeroed the first work counter
cudaError_t launchCheckStates(const SceneDev &sc, const double *dStates, size_t n, uint8_t *dValid,
                              DevCounters *dCtr, unsigned *dWork, int numSms, cudaStream_t stream,
                              bool streamed = false, bool workZeroed = false);

// scratch for the per-warp interval LIFOs of checkEdgesKernel
size_t edgeStackBytes(int numSms);

cudaError_t launchCheckEdges(const SceneDev &sc, const double *dFrom, const double *dTo, const unsigned *dSteps,
                             size_t n, unsigned *dDead, uint8_t *dValid, DevCounters *dCtr, unsigned long long *dWork,
                             void *dStacks, size_t stackBytes, double *dEdgeConst, int numSms, cudaStream_t stream,
                             bool zeroed = false);
// dEdgeConst: edgeConstBytes(n) of scratch for the per-edge interpolation constants
size_t edgeConstBytes(size_t n);
// zeroed: dWork (16 counters) and dDead are already zero.  dValid == nullptr: leave the verdicts as dDead flags.

}  // namespace mplb
