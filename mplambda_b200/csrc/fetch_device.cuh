// Device-side description of a Fetch scene and the launch entry points of fetch_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace mplb {

// fetch_robot.hpp:38-53; FCL conventions ([EXT]): cylinder / capsule axis = local z, centred
struct FetchShape {
    int type;        // 0 box (half sides a,b,c), 1 capsule (radius a, half length b), 2 cylinder (radius a, half length b)
    double a, b, c;
};

struct FetchSceneDev {
    const float4 *envNodes;    // mplb_host.hpp BvhNode, environment-local frame
    const double *envTris64;   // 9 doubles per triangle
    int nEnvNodes;
    double envFrame[12];       // row-major 3x4 [R|t]: environment-local -> world
    double origins[12 * 12];   // link frame -> geometry frame (fetch_robot.hpp:217-274), row-major 3x4 each
    FetchShape shapes[12];
};

struct FetchCounters {
    unsigned long long pairTests, bvTests, leafTests;
    unsigned error, pad;
};

// per-call scratch (sized for the largest launch group of n states)
struct FetchScratch {
    double *frames;              // [n][12][12]
    unsigned char *selfHit;      // [n]
    unsigned *envHit;            // [n]
    unsigned *survivors;         // [n] indices of the states that passed the self-collision tests
    unsigned long long *pairItems, *leafItems;
    size_t pairCap, leafCap;
    unsigned *counts;            // [4]: pair list, leaf list, survivor list lengths, env item counter
};
inline size_t fetchPairCap(size_t n) { return n * 8; }
inline size_t fetchLeafCap(size_t n) { return n * 16; }

cudaError_t launchFetchStates(const FetchSceneDev &sc, const double *dStates, size_t n, const FetchScratch &w,
                              uint8_t *dValid, FetchCounters *dCtr, int numSms, cudaStream_t stream);

// items: (edge << 32) | sample index k; k >= steps[edge] means the edge's `to` state
cudaError_t launchFetchEdgeItems(const FetchSceneDev &sc, const double *dFrom, const double *dTo, const unsigned *dSteps,
                                 const unsigned long long *dItems, size_t nItems, unsigned *dDead, const FetchScratch &w,
                                 FetchCounters *dCtr, int numSms, cudaStream_t stream);

}  // namespace mplb
