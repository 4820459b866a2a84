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
    // Interval certificates for edges: while the joints move by dq, no point of geometry k moves farther than
    // |dq_torso| * lift[k] + sum_{arm joints} |dq_j| * reach[k]  (reach = largest distance from the shoulder pan
    // axis' origin -- hence from any arm joint's axis -- to a point of the geometry; 0 for what the arm does not move)
    float lift[12], reach[12];
};

struct FetchCounters {
    unsigned long long pairTests, bvTests, leafTests;
    unsigned error, pad;   // pad: 1 = a work list overflowed and leaf pairs were dropped: the call must be repeated (fetch_capi.cu)
};

// per-call scratch (sized for the largest launch group of n states)
struct FetchScratch {
    double *frames;              // [n][12][12]
    unsigned char *selfHit;      // [n]
    unsigned *envHit;            // [n]
    unsigned *survivors;         // [n] indices of the states that passed the self-collision tests
    unsigned long long *pairItems, *leafItems;   // pairItems: 2 x pairCap (the pairs that need the exact test follow the list)
    size_t pairCap, leafCap;
    unsigned *counts;            // [12] ([7] redo list length, [8] its item cursor): pair list, leaf list, survivor list lengths, env item counter, exact-pair list
                                 //      length, the two fetchMprKernel item cursors
    unsigned char *touched;      // [n] interval items: something came within reach, the interval is not certified
};
inline size_t fetchPairCap(size_t n) { return n * 8; }
inline size_t fetchLeafCap(size_t n) { return n * 16; }

cudaError_t launchFetchStates(const FetchSceneDev &sc, const double *dStates, size_t n, const FetchScratch &w,
                              uint8_t *dValid, FetchCounters *dCtr, int numSms, cudaStream_t stream);

// items: (edge << 32) | sample index k; k >= steps[edge] means the edge's `to` state.  dSlack (optional, [nItems]):
// the item stands for an interval of samples around k over which the torso joint moves by at most .x and the arm
// joints together by at most .y (radians, L1); w.touched[i] = 1 unless the whole interval is certainly free.
cudaError_t launchFetchEdgeItems(const FetchSceneDev &sc, const double *dFrom, const double *dTo, const unsigned *dSteps,
                                 const unsigned long long *dItems, const float2 *dSlack, size_t nItems, unsigned *dDead,
                                 const FetchScratch &w, FetchCounters *dCtr, int numSms, cudaStream_t stream);

// Interval lists of an edge batch kept on the device (fetch_kernels.cu): per-edge joint motion, the first round's
// intervals (`to` + up to 16 per edge), a group's intervals -> items / slacks / forced flags, and the split of what a
// group could neither certify nor kill into the next round's list.  *dOverflow is set when a list does not hold it all.
cudaError_t launchFetchEdgeMotion(const double *dFrom, const double *dTo, size_t n, float2 *dMotion, cudaStream_t stream);
cudaError_t launchFetchRound0(const unsigned *dSteps, size_t n, uint4 *dList, unsigned *dCount, unsigned cap, unsigned *dOverflow,
                              cudaStream_t stream);
cudaError_t launchFetchItems(const uint4 *dList, size_t cnt, const unsigned *dSteps, const float2 *dMotion, float cap,
                             unsigned long long *dItems, float2 *dSlack, unsigned char *dForced, cudaStream_t stream);
cudaError_t launchFetchSplit(const uint4 *dList, size_t cnt, const unsigned char *dTouched, const unsigned char *dForced,
                             const unsigned *dDead, const unsigned *dSteps, const float2 *dMotion, float cap, unsigned minParts,
                             uint4 *dNext, unsigned *dNextCount, unsigned nextCap, unsigned *dOverflow, cudaStream_t stream);

}  // namespace mplb
