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

cudaError_t launchFetchStates(const FetchSceneDev &sc, const double *dStates, size_t n, double *dFrames,
                              unsigned char *dSelfHit, unsigned *dEnvHit, uint8_t *dValid, FetchCounters *dCtr,
                              cudaStream_t stream);

// items: (edge << 32) | sample index k; k >= steps[edge] means the edge's `to` state
cudaError_t launchFetchEdgeItems(const FetchSceneDev &sc, const double *dFrom, const double *dTo, const unsigned *dSteps,
                                 const unsigned long long *dItems, size_t nItems, unsigned *dDead, double *dFrames,
                                 unsigned char *dSelfHit, unsigned *dEnvHit, FetchCounters *dCtr, cudaStream_t stream);

}  // namespace mplb
