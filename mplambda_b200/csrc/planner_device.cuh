// Launch entry points of the device-resident planner iteration (planner_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace mplb {

// Scenario::randomSample for n SE(3) states (reference include/mpl/demo/se3_rigid_body_scenario.hpp:107-113 with
// include/mpl/randomize.hpp:12-38): state i of the call comes from the counter-based generator Philox4x32-10 keyed by
// `seed` at counter first + i.  goalBias > 0: sample i is replaced by `goal` when its seventh uniform is below the bias
// (prrt.hpp:301-315), dGoalSample[i] says so (may be null).
struct SampleSe3 {
    double lo[3], hi[3], goal[7];
    double goalBias;
    unsigned long long seed, first;
};
cudaError_t launchSampleSe3(const SampleSe3 &p, size_t n, double *dStates, uint8_t *dGoalSample, cudaStream_t stream);

// n states of `dim` (<= 8) coordinates, coordinate j uniform in [lo[j], hi[j]) (randomize.hpp:27-37,
// fetch_robot.hpp:163-175)
struct SampleBox {
    double lo[8], hi[8];
    int dim;
    unsigned long long seed, first;
};
cudaError_t launchSampleBox(const SampleBox &p, size_t n, double *dStates, cudaStream_t stream);

// What one iteration of the device-resident RRT loop reports to the host (pinned memory).
struct RrtStatus {
    unsigned long long added;         // samples that became nodes in this iteration
    long long goalNode;               // lowest new node index that is a goal (-1: none)
    unsigned goalUncertain;           // new nodes whose distance to the goal is within rounding of the goal radius
    unsigned candidates;              // samples with a nearest node at distance > 0 (motions that were asked for)
};

// prrt.hpp:280-299 for a batch: sample i becomes node (count + its rank among the accepted samples) iff its motion from
// the nearest node is valid and that node is not the sample itself.  Appends keys (FP64 + FP32 copy) and parents in
// sample order, so the tree is the one a sequential loop over the batch builds.
struct RrtSelect {
    const double *samples;            // [n][7]
    const uint8_t *goalSample;        // [n]
    const uint8_t *valid;             // [n] verdict of the motion nearest -> sample
    const long long *nearIdx;         // [n]
    const double *nearDist;           // [n]
    double *keys;                     // the nearest-neighbour structure's keys [capacity][7]
    float *keys32;
    long long *parent;                // [capacity]
    unsigned long long count;         // nodes so far
    double goal[7], goalRadius, so3w, l2w;
};
cudaError_t launchRrtSelect(const RrtSelect &p, size_t n, RrtStatus *dStatus, cudaStream_t stream);

}  // namespace mplb
