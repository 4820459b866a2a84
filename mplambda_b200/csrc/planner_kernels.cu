// Kernels of the device-resident planner iteration: the sampler (Scenario::randomSample) and the step that turns the
// verdicts of a batch into new tree nodes.  Together with the nearest-neighbour search (nn_kernels.cu) and the
// index-named motion check (se3_kernels.cu: edgeGatherPlanKernel + checkEdgesKernel) one iteration of
// Planner<Scenario, PRRT>::Thread::addSample (reference include/mpl/prrt.hpp:280-299) runs without a state crossing
// PCIe: the host sees a 24-byte status per iteration.
#include "planner_device.cuh"
#include "device_math.cuh"

namespace mplb {
namespace {

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"), counter (c0, c1, 0, 0), key (k0, 0):
// the generator of mplambda_b200/workload.py, so host and device name the same uniforms.
__device__ __forceinline__ void philox4x32(unsigned c0, unsigned c1, unsigned k0, unsigned out[4]) {
    unsigned c2 = 0, c3 = 0, k1 = 0;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u01(unsigned x) { return dmul((double)x, 1.0 / 4294967296.0); }

__global__ void sampleSe3Kernel(SampleSe3 p, unsigned long long n, double *__restrict__ states,
                                uint8_t *__restrict__ goalSample) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long ctr = p.first + i;
    unsigned r0[4], r1[4];
    philox4x32((unsigned)ctr, 0u, (unsigned)p.seed, r0);   // (workload.py: counter (i, call, 0, 0), i taken mod 2^32)
    philox4x32((unsigned)ctr, 1u, (unsigned)p.seed, r1);
    double *o = states + 7 * i;
    const bool isGoal = p.goalBias > 0.0 && u01(r1[2]) < p.goalBias;
    if (goalSample) goalSample[i] = isGoal;
    if (isGoal) {
#pragma unroll
        for (int k = 0; k < 7; ++k) o[k] = p.goal[k];
        return;
    }
    // randomize.hpp:16-24
    const double a = u01(r0[0]), b = dmul(6.283185307179586, u01(r0[1])), c = dmul(6.283185307179586, u01(r0[2]));
    const double sa = sqrt(dsub(1.0, a)), sb = sqrt(a);
    double sinB, cosB, sinC, cosC;
    sincos(b, &sinB, &cosB);
    sincos(c, &sinC, &cosC);
    o[0] = dmul(sa, sinB);
    o[1] = dmul(sa, cosB);
    o[2] = dmul(sb, sinC);
    o[3] = dmul(sb, cosC);
    // randomize.hpp:33-36 (uniform_real_distribution: min + u (max - min))
    o[4] = dadd(p.lo[0], dmul(u01(r0[3]), dsub(p.hi[0], p.lo[0])));
    o[5] = dadd(p.lo[1], dmul(u01(r1[0]), dsub(p.hi[1], p.lo[1])));
    o[6] = dadd(p.lo[2], dmul(u01(r1[1]), dsub(p.hi[2], p.lo[2])));
}

__global__ void sampleBoxKernel(SampleBox p, unsigned long long n, double *__restrict__ states) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long ctr = p.first + i;
    unsigned r[8];
    philox4x32((unsigned)ctr, 0u, (unsigned)p.seed, r);
    philox4x32((unsigned)ctr, 1u, (unsigned)p.seed, r + 4);
    for (int k = 0; k < p.dim; ++k) states[p.dim * i + k] = dadd(p.lo[k], dmul(u01(r[k]), dsub(p.hi[k], p.lo[k])));
}

// One CTA walks the batch in sample order: a block-wide exclusive scan of the accepted flags gives every accepted
// sample its node index, so the result does not depend on how the work is split.
constexpr int kSelectThreads = 1024;
__global__ void __launch_bounds__(kSelectThreads) rrtSelectKernel(RrtSelect p, unsigned long long n, RrtStatus *status) {
    __shared__ unsigned warpSum[32];
    __shared__ unsigned long long base;
    __shared__ unsigned cand, uncertain;
    __shared__ long long goalNode;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        base = p.count;
        cand = uncertain = 0;
        goalNode = 0x7fffffffffffffffll;
    }
    __syncthreads();
    for (unsigned long long lo = 0; lo < n; lo += kSelectThreads) {
        const unsigned long long i = lo + threadIdx.x;
        bool isCand = false, ok = false;
        if (i < n) {
            isCand = p.nearDist[i] > 0.0 && p.nearIdx[i] >= 0;   // prrt.hpp:283-284
            ok = isCand && p.valid[i] != 0;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, ok), cballot = __ballot_sync(0xffffffffu, isCand);
        if (lane == 0) warpSum[warp] = __popc(ballot);
        if (lane == 0 && cballot) atomicAdd(&cand, (unsigned)__popc(cballot));
        __syncthreads();
        unsigned before = 0, total = 0;
        for (unsigned w = 0; w < kSelectThreads / 32; ++w) {
            const unsigned s = warpSum[w];
            before += w < warp ? s : 0u;
            total += s;
        }
        if (ok) {
            const unsigned long long node = base + before + __popc(ballot & ((1u << lane) - 1u));
            const double *q = p.samples + 7 * i;
            double *k = p.keys + 7 * node;
            float *k32 = p.keys32 + 7 * node;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                k[c] = q[c];
                k32[c] = (float)q[c];
            }
            p.parent[node] = p.nearIdx[i];
            // isGoal (se3_rigid_body_scenario.hpp:115-117): distance(goal, q) <= goalRadius; sampled goals are goals
            bool goal = p.goalSample && p.goalSample[i];
            if (!goal) {
                const double d = fabs(dadd(dadd(dadd(dmul(p.goal[0], q[0]), dmul(p.goal[1], q[1])), dmul(p.goal[2], q[2])),
                                           dmul(p.goal[3], q[3])));
                const double so3 = d < 1.0 ? acos(d) : 0.0;
                const double dx = dsub(p.goal[4], q[4]), dy = dsub(p.goal[5], q[5]), dz = dsub(p.goal[6], q[6]);
                const double dist = dadd(dmul(p.so3w, so3), dmul(p.l2w, sqrt(dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz)))));
                goal = dist <= p.goalRadius;
                // CUDA's acos is not glibc's: a distance this close to the radius is decided by the host
                if (fabs(dist - p.goalRadius) <= dmul(p.so3w, 2e-15) + dist * 1e-15) atomicAdd(&uncertain, 1u);
            }
            if (goal) atomicMin((unsigned long long *)&goalNode, node);
        }
        __syncthreads();
        if (threadIdx.x == 0) base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        status->added = base - p.count;
        status->goalNode = goalNode == 0x7fffffffffffffffll ? -1ll : goalNode;
        status->goalUncertain = uncertain;
        status->candidates = cand;
    }
}

}  // namespace

cudaError_t launchSampleSe3(const SampleSe3 &p, size_t n, double *dStates, uint8_t *dGoalSample, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    sampleSe3Kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(p, n, dStates, dGoalSample);
    return cudaGetLastError();
}

cudaError_t launchSampleBox(const SampleBox &p, size_t n, double *dStates, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    sampleBoxKernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(p, n, dStates);
    return cudaGetLastError();
}

cudaError_t launchRrtSelect(const RrtSelect &p, size_t n, RrtStatus *dStatus, cudaStream_t stream) {
    rrtSelectKernel<<<1, kSelectThreads, 0, stream>>>(p, n, dStatus);
    return cudaGetLastError();
}

}  // namespace mplb
