// SE(3) rigid-body validity kernels for sm_100a.
//
// Device replacement for the work behind
//   SE3RigidBodyScenario::isValid(q)        reference include/mpl/demo/se3_rigid_body_scenario.hpp:119-126
//   SE3RigidBodyScenario::isValid(from,to)  reference include/mpl/demo/se3_rigid_body_scenario.hpp:134-178
//   mpl::interpolate (slerp / lerp)         reference include/mpl/interpolate.hpp:9-36
// i.e. fcl::collide(robotMesh, tf(q), envMesh, Identity) per state ([EXT] FCL: OBB hierarchy
// traversal + 17-axis triangle-triangle separating-axis test, first contact wins).
//
// Execution model (one warp = one streaming work pool over 64 state slots in shared memory):
//   * states are expanded 32 at a time (FP64 quaternion/slerp math -> FP32 (R|T) slot) into free
//     slots; a lane that runs out of work adopts the next ready slot, so all 32 lanes stay busy
//     however uneven the per-state traversal depth is, and every state is still searched
//     depth-first (first contact is found after few box tests, as in the reference's recursion);
//   * each lane owns a private LIFO of (robotNode, envNode) pairs in shared memory.  One step
//     pops a pair known to overlap, loads the two children of the larger box and runs both
//     OBB-OBB separating-axis tests; survivors are pushed nearest-centre-first;
//   * leaf-leaf survivors go to a warp-shared queue and are tested 32 at a time with an
//     error-bounded FP32 17-axis SAT; the rare pairs whose FP32 verdict is not certain go to a
//     third queue and are re-evaluated in FP64 with exactly the oracle's operation order
//     (__dmul_rn/__dadd_rn: no FMA contraction), which is what makes verdicts bit-identical to
//     the FP64 CPU path instead of "equal up to an epsilon band";
//   * hit/ready/free slot sets are warp-uniform bit masks maintained with ballot/reduce_or; a
//     slot is recycled once its search is over and none of its leaf pairs is still queued.
//   * lanes with nothing left steal the bottom pair of a busy lane (the largest pending subtree);
//   * work is handed out by an atomic counter in guided chunks, so the end of a batch -- and all
//     of a small one -- spreads over the whole grid; edges are handed out as intervals of samples
//     that are certified free as a whole where possible (EdgeSource), in three passes.
// Node fetches are 256-bit read-only loads (2 per 64-byte node; sm_100's LDG.E.256), triangles
// one 256-bit and one 128-bit load: a scattered load costs one L1 wavefront per lane and
// instruction whatever its width, and the L1 data pipe is what binds the box-pair step.  The
// hierarchy lives in L1/L2, HBM only sees states in and verdict bytes out.
#include "se3_device.cuh"
#include "device_math.cuh"

#include <algorithm>
#include <cstdlib>
#include <cfloat>
#include <map>
#include <mutex>
#include <tuple>

namespace mplb {
namespace {

#ifndef MPLB_CLAIM_MIN
#define MPLB_CLAIM_MIN 4
#endif
// One CTA per SM, 12 warps: the kernels need ~170 registers to run without spilling their warp state (slot masks, work
// source, counters) to local memory, and 12 x 32 x 170 is what a register file holds.  Measured against 2 CTAs x 8 warps
// at 128 registers (2^20 poses): Alpha 0.794 -> 0.744 ms, Cubicles 0.99 -> 0.95, Apartment 4.18 -> 3.40; 65,536 Cubicles
// edges 0.93 -> 0.79 ms.  (6 x 2, 10, 11, 14 and 16 warps were measured too; 14 and 16 no longer fit Apartment's stacks.)
#ifndef MPLB_WARPS_PER_CTA
#define MPLB_WARPS_PER_CTA 12
#endif
#ifndef MPLB_EDGE_WARPS
#define MPLB_EDGE_WARPS 12
#endif
#ifndef MPLB_MIN_CTAS
#define MPLB_MIN_CTAS 1
#endif
#ifndef MPLB_WARPS_SMALL_CTA
#define MPLB_WARPS_SMALL_CTA MPLB_WARPS_PER_CTA   // (4: see launchStatesT -- measured, not enabled)
#endif
constexpr int kWarpsStates = MPLB_WARPS_PER_CTA, kWarpsEdges = MPLB_EDGE_WARPS;
constexpr int kWarpsSmall = MPLB_WARPS_SMALL_CTA;   // pose batches: CTAs of this many warps where three of them fit an SM (launchStatesT)
#ifndef MPLB_SLOTS
#define MPLB_SLOTS 64
#endif
constexpr int kSlots = MPLB_SLOTS;  // state slots per warp (64 or 32)
constexpr int kQueueCap = 96;       // leaf / exact queue entries per warp
#ifndef MPLB_EDGE_INTERVAL_BITS
#define MPLB_EDGE_INTERVAL_BITS 5
#endif
constexpr unsigned kIntervalBits = MPLB_EDGE_INTERVAL_BITS;
constexpr unsigned kPhaseBSub = 1u << kIntervalBits;    // intervals an edge's samples are cut into (see EdgeSource)
#ifndef MPLB_SPILL_KEEP
#define MPLB_SPILL_KEEP 0
#endif
constexpr unsigned kSpillKeep = MPLB_SPILL_KEEP;    // interval pieces a warp keeps for itself once the item pool has drained
constexpr size_t kLatencyBelow = 1u << 17;   // poses: below this checkStatesKernel's latency variant is launched
#ifndef MPLB_STEPS_PER_ROUND
#define MPLB_STEPS_PER_ROUND 4
#endif
constexpr int kStepsPerRound = MPLB_STEPS_PER_ROUND;   // depth-first steps between two rounds of slot bookkeeping
constexpr unsigned kFull = 0xffffffffu;

// Per-warp shared memory, carved from the dynamic allocation at run time: the LIFO only gets as
// many levels as the scene's hierarchies need (depthA + depthB + 2), which keeps the footprint
// small enough for L1 to hold the hot part of the hierarchy.
template <typename E>
struct WarpSmem {
    unsigned (*stack)[32];            // [level][lane]: robotIdx | envIdx << bitsA | descendRobot << 31  (bank = lane)
    float4 (*xf)[5];                  // per slot: rows of (R | T), then (slack, eta, -, -); see makeSlot.  80-byte
                                      // stride: lanes reading row r of 32 different slots spread over all banks
                                      // (a 64-byte stride puts them on two 16-byte bank groups: 16-way conflicts)
    unsigned *tag0, *tag1, *tag2, *tag3;  // what the slot is (state index; or edge, sample, interval lo, hi)
    int *pending;                     // leaf pairs of the slot still sitting in triq/exq
    int *workers;                     // lanes currently searching the slot (work stealing splits a search)
    unsigned *list;                   // scratch: compacted slot / lane lists
    uint4 *work;                      // scratch: up to 32 work descriptors of the current refill
    E *triq;                          // slot | robotTri << 6 | envTri << (6 + bitsA)
    E *exq;

    static __host__ __device__ size_t bytes(int levels) {
        return (size_t)levels * 128 + kSlots * 80 + 32 * 16 + kSlots * 28 + 2 * kQueueCap * sizeof(E);
    }
    __device__ void carve(unsigned char *base, int levels) {
        stack = reinterpret_cast<unsigned(*)[32]>(base);
        base += (size_t)levels * 128;
        xf = reinterpret_cast<float4(*)[5]>(base);
        base += kSlots * 80;
        work = reinterpret_cast<uint4 *>(base);
        base += 32 * 16;
        triq = reinterpret_cast<E *>(base);
        base += kQueueCap * sizeof(E);
        exq = reinterpret_cast<E *>(base);
        base += kQueueCap * sizeof(E);
        tag0 = reinterpret_cast<unsigned *>(base);
        tag1 = tag0 + kSlots;
        tag2 = tag1 + kSlots;
        tag3 = tag2 + kSlots;
        pending = reinterpret_cast<int *>(tag3 + kSlots);
        workers = pending + kSlots;
        list = reinterpret_cast<unsigned *>(workers + kSlots);
    }
};

struct XformD { double R[3][3]; double T[3]; };

// Eigen toRotationMatrix order, then R = R1^T, T = R1^T (0 - t1)
__device__ void stateToRelD(const double s[7], XformD &X) {
    double x = s[0], y = s[1], z = s[2], w = s[3];
    double tx = dmul(2, x), ty = dmul(2, y), tz = dmul(2, z);
    double twx = dmul(tx, w), twy = dmul(ty, w), twz = dmul(tz, w);
    double txx = dmul(tx, x), txy = dmul(ty, x), txz = dmul(tz, x);
    double tyy = dmul(ty, y), tyz = dmul(tz, y), tzz = dmul(tz, z);
    double R1[3][3];
    R1[0][0] = dsub(1, dadd(tyy, tzz)); R1[0][1] = dsub(txy, twz);          R1[0][2] = dadd(txz, twy);
    R1[1][0] = dadd(txy, twz);          R1[1][1] = dsub(1, dadd(txx, tzz)); R1[1][2] = dsub(tyz, twx);
    R1[2][0] = dsub(txz, twy);          R1[2][1] = dadd(tyz, twx);          R1[2][2] = dsub(1, dadd(txx, tyy));
    double nt[3] = {dsub(0.0, s[4]), dsub(0.0, s[5]), dsub(0.0, s[6])};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) X.R[i][j] = R1[j][i];
        X.T[i] = dadd(dadd(dmul(R1[0][i], nt[0]), dmul(R1[1][i], nt[1])), dmul(R1[2][i], nt[2]));
    }
}

__device__ __forceinline__ D3 xfApplyD(const XformD &X, D3 p) {
    return {dadd(dadd(dadd(dmul(X.R[0][0], p.x), dmul(X.R[0][1], p.y)), dmul(X.R[0][2], p.z)), X.T[0]),
            dadd(dadd(dadd(dmul(X.R[1][0], p.x), dmul(X.R[1][1], p.y)), dmul(X.R[1][2], p.z)), X.T[1]),
            dadd(dadd(dadd(dmul(X.R[2][0], p.x), dmul(X.R[2][1], p.y)), dmul(X.R[2][2], p.z)), X.T[2])};
}

__device__ __forceinline__ bool project6D(D3 ax, D3 p1, D3 p2, D3 p3, D3 q1, D3 q2, D3 q3) {
    double P1 = d3dot(ax, p1), P2 = d3dot(ax, p2), P3 = d3dot(ax, p3);
    double Q1 = d3dot(ax, q1), Q2 = d3dot(ax, q2), Q3 = d3dot(ax, q3);
    double mx1 = fmax(P1, fmax(P2, P3)), mn1 = fmin(P1, fmin(P2, P3));
    double mx2 = fmax(Q1, fmax(Q2, Q3)), mn2 = fmin(Q1, fmin(Q2, Q3));
    if (mn1 > mx2) return false;
    if (mn2 > mx1) return false;
    return true;
}

// [EXT] fcl Intersect::intersect_Triangle, axis order n1, m1, e_i x f_j, e_i x n1, f_i x m1
__device__ __noinline__ bool triTriExact(const double *P, const double *Q, const XformD &X) {
    D3 P1 = {P[0], P[1], P[2]}, P2 = {P[3], P[4], P[5]}, P3 = {P[6], P[7], P[8]};
    D3 Q1 = xfApplyD(X, {Q[0], Q[1], Q[2]}), Q2 = xfApplyD(X, {Q[3], Q[4], Q[5]}),
       Q3 = xfApplyD(X, {Q[6], Q[7], Q[8]});
    D3 p1 = d3sub(P1, P1), p2 = d3sub(P2, P1), p3 = d3sub(P3, P1);
    D3 q1 = d3sub(Q1, P1), q2 = d3sub(Q2, P1), q3 = d3sub(Q3, P1);
    D3 e[3] = {d3sub(p2, p1), d3sub(p3, p2), d3sub(p1, p3)};
    D3 f[3] = {d3sub(q2, q1), d3sub(q3, q2), d3sub(q1, q3)};
    D3 n1 = d3cross(e[0], e[1]), m1 = d3cross(f[0], f[1]);
    if (!project6D(n1, p1, p2, p3, q1, q2, q3)) return false;
    if (!project6D(m1, p1, p2, p3, q1, q2, q3)) return false;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (!project6D(d3cross(e[i], f[j]), p1, p2, p3, q1, q2, q3)) return false;
    for (int i = 0; i < 3; ++i)
        if (!project6D(d3cross(e[i], n1), p1, p2, p3, q1, q2, q3)) return false;
    for (int i = 0; i < 3; ++i)
        if (!project6D(d3cross(f[i], m1), p1, p2, p3, q1, q2, q3)) return false;
    return true;
}

// Diagnostic twin of triTriExact (SceneDev::trackMargins): all 17 axes, no early exit, and besides the verdict the
// decision's margin -- G = the largest gap any axis leaves between the two projections, relative to |axis|_1 times the
// largest coordinate (the scale rounding errors of the coordinates enter with).  The pair is disjoint iff G > 0; |G|
// says how far the inputs were from the other verdict.  CUDA's sin/acos (edge samples) differ from glibc's by <= 2 ulp,
// ~1e-16 relative: a verdict can only depend on that when |G| is of that order.
__device__ __noinline__ bool triTriExactMargin(const double *P, const double *Q, const XformD &X, float &margin) {
    D3 P1 = {P[0], P[1], P[2]}, P2 = {P[3], P[4], P[5]}, P3 = {P[6], P[7], P[8]};
    D3 Q1 = xfApplyD(X, {Q[0], Q[1], Q[2]}), Q2 = xfApplyD(X, {Q[3], Q[4], Q[5]}), Q3 = xfApplyD(X, {Q[6], Q[7], Q[8]});
    D3 p1 = d3sub(P1, P1), p2 = d3sub(P2, P1), p3 = d3sub(P3, P1);
    D3 q1 = d3sub(Q1, P1), q2 = d3sub(Q2, P1), q3 = d3sub(Q3, P1);
    D3 e[3] = {d3sub(p2, p1), d3sub(p3, p2), d3sub(p1, p3)};
    D3 f[3] = {d3sub(q2, q1), d3sub(q3, q2), d3sub(q1, q3)};
    D3 n1 = d3cross(e[0], e[1]), m1 = d3cross(f[0], f[1]);
    auto amax = [](D3 a) { return fmax(fabs(a.x), fmax(fabs(a.y), fabs(a.z))); };
    const double U = fmax(fmax(fmax(amax(p2), amax(p3)), fmax(amax(q1), amax(q2))), amax(q3));
    double G = -1e300;
    bool sep = false;
    auto axis = [&](D3 ax) {
        const double a1 = fabs(ax.x) + fabs(ax.y) + fabs(ax.z);
        double Pv1 = d3dot(ax, p1), Pv2 = d3dot(ax, p2), Pv3 = d3dot(ax, p3);
        double Qv1 = d3dot(ax, q1), Qv2 = d3dot(ax, q2), Qv3 = d3dot(ax, q3);
        double mx1 = fmax(Pv1, fmax(Pv2, Pv3)), mn1 = fmin(Pv1, fmin(Pv2, Pv3));
        double mx2 = fmax(Qv1, fmax(Qv2, Qv3)), mn2 = fmin(Qv1, fmin(Qv2, Qv3));
        if (mn1 > mx2 || mn2 > mx1) sep = true;   // the reference's comparisons, unchanged
        if (a1 > 0 && U > 0) G = fmax(G, fmax(mn1 - mx2, mn2 - mx1) / (a1 * U));
    };
    axis(n1);
    axis(m1);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) axis(d3cross(e[i], f[j]));
    for (int i = 0; i < 3; ++i) axis(d3cross(e[i], n1));
    for (int i = 0; i < 3; ++i) axis(d3cross(f[i], m1));
    margin = G > -1e299 ? (float)fabs(G) : INFINITY;
    return !sep;
}

// The exact re-evaluation of one triangle pair, out of line: the state is rebuilt from the slot's tags by the source's
// fetcher (a few pointers, passed by value) -- the FP64 code stays out of the hot loop's instruction footprint.
template <typename Fetch>
__device__ __noinline__ bool exactPairHit(const Fetch fetch, unsigned t0, unsigned t1, const double *P, const double *Q) {
    double s[7];
    fetch.stateOf(t0, t1, s);
    XformD X;
    stateToRelD(s, X);
    return triTriExact(P, Q, X);
}
// the same with the decision's margin recorded (diagnostic; a function of its own so that the one above -- and the
// kernel around its call -- compile exactly as they do without the diagnostic: an extra argument there cost 25 %)
template <typename Fetch>
__device__ __noinline__ bool exactPairHitMargin(const Fetch fetch, unsigned t0, unsigned t1, const double *P, const double *Q,
                                                unsigned *minMargin) {
    double s[7];
    fetch.stateOf(t0, t1, s);
    XformD X;
    stateToRelD(s, X);
    float margin;
    const bool hit = triTriExactMargin(P, Q, X, margin);
    // smallest margin so far, kept as ~bits so that the zeroed counter block means "none yet"
    atomicMax(minMargin, ~__float_as_uint(margin));
    return hit;
}

// interpolate.hpp:18-36 with Eigen's slerp; same statements as orc_se3_interpolate.  sin/acos are
// CUDA's FP64 libdevice (<= 2 ulp from glibc's, see DESIGN.md "numerics").  What depends on the edge only
// (d, theta = acos|d|, sin theta; and the two terms of the interval motion bound) is computed once per
// edge by edgeConstKernel -- the same operations, hoisted: the per-sample conversion is the code every
// warp of checkEdgesKernel keeps coming back to, and that kernel is bound by instruction fetch.
constexpr int kEdgeConst = 5;   // doubles per edge: d, theta (< 0: linear blend), sin theta, motion per unit reach, scale term

__device__ __forceinline__ void interpolateWith(const double *from, const double *to, double t, double d, double theta,
                                                double sinTheta, double out[7]) {
    double s0, s1;
    if (theta < 0) {
        s0 = dsub(1.0, t);
        s1 = t;
    } else {
        s0 = __ddiv_rn(sin(dmul(dsub(1.0, t), theta)), sinTheta);
        s1 = __ddiv_rn(sin(dmul(t, theta)), sinTheta);
    }
    if (d < 0) s1 = -s1;
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = dadd(dmul(s0, from[i]), dmul(s1, to[i]));
#pragma unroll
    for (int i = 4; i < 7; ++i) out[i] = dadd(from[i], dmul(dsub(to[i], from[i]), t));
}

// per-edge constants of edge (a -> b) into o[kEdgeConst]
__device__ __forceinline__ void edgeConstOf(const double *a, const double *b, float robotRadius, double *o) {
    const double d = dadd(dadd(dadd(dmul(a[0], b[0]), dmul(a[1], b[1])), dmul(a[2], b[2])), dmul(a[3], b[3]));
    const double absD = fabs(d);
    double theta = -1.0, sinTheta = 0.0;
    if (!(absD >= 1.0 - DBL_EPSILON)) {
        theta = acos(absD);
        sinTheta = sin(theta);
    }
    // Motion bound (EdgeSource::load): translation is a lerp and Eigen's slerp turns at constant rate, so over a
    // fraction `reach` of the edge a robot point x moves at most reach * (||dt|| + 2 theta |x|)
    double na = 0, nb = 0, dt2 = 0;
    for (int k = 0; k < 4; ++k) {
        na += a[k] * a[k];
        nb += b[k] * b[k];
    }
    for (int k = 4; k < 7; ++k) {
        const double x = b[k] - a[k];
        dt2 += x * x;
    }
    // rotation matrices of non-unit quaternions scale: the bound assumes rigid motion
    const double scale = fmax(1.0, fmax(na, nb));
    const double thetaM = acos(fmin(1.0, absD / sqrt(na * nb)));
    const double R = (double)robotRadius * scale;
    o[0] = d;
    o[1] = theta;
    o[2] = sinTheta;
    o[3] = sqrt(dt2) + 2.0 * thetaM * R;
    o[4] = fabs(scale - 1.0) * 4.0 * R * 1.000001 + 1e-6 * R;
}

__global__ void edgeConstKernel(const double *__restrict__ from, const double *__restrict__ to, unsigned long long n,
                                float robotRadius, double *__restrict__ out) {
    const unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (e >= n) return;
    edgeConstOf(from + 7 * e, to + 7 * e, robotRadius, out + kEdgeConst * e);
}

// Direct-return host calls: the end states and host-planned step counts arrive in mapped host memory; they are read
// across the link once, here, and left in device memory for the passes (which read an edge's ends once per sample).
__global__ void edgeStageKernel(const double *__restrict__ hostIn, unsigned long long n, float robotRadius,
                                double *__restrict__ from, double *__restrict__ to, unsigned *__restrict__ steps,
                                double *__restrict__ out) {
    const unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (e >= n) return;
    double a[7], b[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        a[k] = hostIn[7 * e + k];
        b[k] = hostIn[7 * (n + e) + k];
        from[7 * e + k] = a[k];
        to[7 * e + k] = b[k];
    }
    steps[e] = reinterpret_cast<const unsigned *>(hostIn + 14 * n)[e];
    edgeConstOf(a, b, robotRadius, out + kEdgeConst * e);
}

// Edges named by index (the planner's "node i of the tree -> sample j" without the states crossing PCIe): gathers both
// end states into dense arrays the edge kernel reads, and plans the edge on the device --
//   steps = ceil(distance(from, to) * invStepSize)        (se3_rigid_body_scenario.hpp:139, nigh SE3Space distance)
// in the host path's operation order.  Every operation but acos is correctly rounded on both sides; CUDA's acos may
// differ from glibc's in the last bits, which changes `steps` only when distance * invStep lies within a few ulp of an
// integer: such edges are counted and listed (DevCounters::ambiguous) so that the caller can plan them with host libm --
// they are never silently different.  An index < 0 or past its pool, or a non-finite distance, kills the edge.
__global__ void edgeGatherPlanKernel(EdgeGather g, unsigned long long n, double so3w, double l2w, double invStep,
                                     float robotRadius, double *__restrict__ from, double *__restrict__ to,
                                     unsigned *__restrict__ steps, unsigned *__restrict__ dead,
                                     double *__restrict__ edgeConst, DevCounters *ctr) {
    const unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (e >= n) return;
    const long long ia = g.fromIdx ? g.fromIdx[e] : (long long)(e / g.fromDiv);
    const long long ib = g.toIdx ? g.toIdx[e] : (long long)(e / g.toDiv);
    double a[7], b[7];
    const bool okIdx = ia >= 0 && ib >= 0 && (unsigned long long)ia < g.fromRows && (unsigned long long)ib < g.toRows;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        a[k] = okIdx ? g.fromPool[7 * ia + k] : 0.0;
        b[k] = okIdx ? g.toPool[7 * ib + k] : 0.0;
        from[7 * e + k] = a[k];
        to[7 * e + k] = b[k];
    }
    if (!okIdx) {
        steps[e] = 0;
        dead[e] = 1u;
#pragma unroll
        for (int k = 0; k < kEdgeConst; ++k) edgeConst[kEdgeConst * e + k] = 0.0;
        return;
    }
    edgeConstOf(a, b, robotRadius, edgeConst + kEdgeConst * e);
    const double d = fabs(edgeConst[kEdgeConst * e]);
    const double so3 = d < 1.0 ? acos(d) : 0.0;
    const double dx = dsub(a[4], b[4]), dy = dsub(a[5], b[5]), dz = dsub(a[6], b[6]);
    const double l2 = sqrt(dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz)));
    const double x = dmul(dadd(dmul(so3w, so3), dmul(l2w, l2)), invStep);
    if (!(x < 2147483648.0)) {   // non-finite state: the host path rejects the batch, here the edge is invalid
        steps[e] = 0;
        dead[e] = 1u;
        atomicAdd(&ctr->badEdges, 1u);
        return;
    }
    steps[e] = (unsigned)ceil(x);
    // |x_device - x_host| <= so3w * invStep * (acos: 2 + 1 ulp of pi/2) + rounding of the two products and the sum
    const bool near = d < 1.0 && fabs(x - rint(x)) <= dmul(dmul(so3w, invStep), 2e-15) + x * 2e-15;
    if (near || (g.forceAmbiguousEvery && e % g.forceAmbiguousEvery == 0)) {
        const unsigned at = atomicAdd(&ctr->ambiguous, 1u);
        if (at < kAmbiguousListed) ctr->ambiguousEdge[at] = (unsigned)e;
    }
}

// ---------------------------------------------------------------------------------------------
// FP32 fast path

// Round the FP64 relative transform to the FP32 slot and derive the error parameters.
// extraSlack > 0 turns the slot into an interval certificate: every box test is inflated by the largest
// displacement any robot point can have, relative to this pose, over the interval the slot stands for.
__device__ __forceinline__ void makeSlot(const double s[7], float envRadius, float robotRadius, float4 *xf, float extraSlack = 0.f) {
    XformD X;
    stateToRelD(s, X);
    double qn = s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + s[3] * s[3];
    double scale = fmax(1.0, qn);  // |R| for non-unit quaternions
    double tn = sqrt(s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
    double mag = scale * ((double)envRadius + tn) + (double)robotRadius;
    float eta = (float)(8.0 * FLT_EPSILON * mag) * 1.0000002f;
    float slack = (float)(64.0 * FLT_EPSILON * (mag + 4.0 * scale * (double)envRadius +
                                                4.0 * (double)robotRadius));
    xf[0] = make_float4((float)X.R[0][0], (float)X.R[0][1], (float)X.R[0][2], (float)X.T[0]);
    xf[1] = make_float4((float)X.R[1][0], (float)X.R[1][1], (float)X.R[1][2], (float)X.T[1]);
    xf[2] = make_float4((float)X.R[2][0], (float)X.R[2][1], (float)X.R[2][2], (float)X.T[2]);
    xf[3] = make_float4(slack + extraSlack, eta, slack, 0.f);  // .z keeps the plain slack
}

// One new slot: the source's loader -- a few pointers, by value -- rebuilds the state of work item `rank` (`work`: its
// descriptor, for sources that use one), the slot's transform and tags are written to shared memory (tags: tag0 +
// slot; tag1..3 and pending follow at kSlots strides, see WarpSmem).  -> false: the item needs no work.
// (Out of line it was measured slower, 0.786 -> 0.805 ms per 2^20 Alpha poses: the call's register traffic costs more
// than the ~300 FP64 instructions take from the instruction cache.)
template <typename Loader>
__device__ __forceinline__ bool fillSlotBody(const Loader ld, const uint4 work, int rank, float envRadius, float robotRadius,
                                             float4 *xf, unsigned *tags) {
    double s[7];
    unsigned t[4];
    float extraSlack;
    if (!ld.load(rank, work, s, t, extraSlack)) return false;
    makeSlot(s, envRadius, robotRadius, xf, extraSlack);
    tags[0] = t[0];
    tags[kSlots] = t[1];
    tags[2 * kSlots] = t[2];
    tags[3 * kSlots] = t[3];
    tags[4 * kSlots] = 0;   // pending
    return true;
}

#ifndef MPLB_EDGE_FILL_OUTLINE
#define MPLB_EDGE_FILL_OUTLINE 0
#endif
template <typename Loader>
__device__ __noinline__ bool fillSlotOutOfLine(const Loader ld, const uint4 work, int rank, float envRadius, float robotRadius,
                                               float4 *xf, unsigned *tags) {
    return fillSlotBody(ld, work, rank, envRadius, robotRadius, xf, tags);
}
template <bool kOutline, typename Loader>
__device__ __forceinline__ bool fillSlot(const Loader ld, const uint4 work, int rank, float envRadius, float robotRadius,
                                         float4 *xf, unsigned *tags) {
    if constexpr (kOutline) return fillSlotOutOfLine(ld, work, rank, envRadius, robotRadius, xf, tags);
    else return fillSlotBody(ld, work, rank, envRadius, robotRadius, xf, tags);
}

enum TriVerdict : int { kTriNo = 0, kTriYes = 1, kTriUnsure = 2 };

struct AxisState {
    bool sep;     // some axis separates beyond its error bound -> certainly disjoint
    bool unsure;  // some axis is within its error bound of separating
};

__device__ __forceinline__ float absMax(F3 a) { return fmaxf(fabsf(a.x), fmaxf(fabsf(a.y), fabsf(a.z))); }

// One separating axis.  E = |ax|_1 * S + K bounds |gap32 - gap64| (see triTriFilter).
__device__ __forceinline__ void axisTest(F3 ax, F3 p2, F3 p3, F3 q1, F3 q2, F3 q3, float S, float K, AxisState &st) {
    // p1 is the origin of the relative frame: its projection is exactly 0
    const float P2 = dot(ax, p2), P3 = dot(ax, p3);
    const float Q1 = dot(ax, q1), Q2 = dot(ax, q2), Q3 = dot(ax, q3);
    const float mx1 = fmaxf(0.f, fmaxf(P2, P3)), mn1 = fminf(0.f, fminf(P2, P3));
    const float mx2 = fmaxf(Q1, fmaxf(Q2, Q3)), mn2 = fminf(Q1, fminf(Q2, Q3));
    const float g = fmaxf(mn1 - mx2, mn2 - mx1);  // > 0 <=> this axis separates
    const float E = fmaf(fabsf(ax.x) + fabsf(ax.y) + fabsf(ax.z), S, K);
    st.sep |= g > E;
    st.unsure |= g >= -E;
}

// Error-bounded FP32 version of the 17-axis test (FCL axis order).  Inputs: robot triangle as
// (P1, p2 = P2-P1, p3 = P3-P1) and environment triangle as (Q1, F1 = Q2-Q1, F2 = Q3-Q1), the
// differences formed in FP64 on the host, so that only q1 = R*Q1 + T - P1 carries the large
// cancellation error eta; edge vectors only carry relative rounding.  With
//   Me, Mf = max |component| of the robot / environment edge vectors, U = max |relative coord|,
// the FP32 gap of an axis differs from the FP64 specification's gap by at most
//   E = |ax|_1 * (eta + 79 eps U) + 6 U * (abs. error of the axis components)
// where the axis error is 20 eps Me^2 (n1), 132 eps Mf^2 (m1), 76 eps Me Mf (e x f),
// 64 eps Me^3 (e x n1), 400 eps Mf^3 (f x m1); derivation in DESIGN.md "FP32 filter", numerical
// check in tests/test_filter_bound.py.  eps = FLT_EPSILON = 2 * unit roundoff (2x safety).
// Certain answers need no FP64 work; kTriUnsure pairs are re-evaluated exactly.
__device__ __forceinline__ int triTriFilter(F3 p2, F3 p3, F3 q1, F3 f1, F3 f2, float eta) {
    const F3 q2 = f3(q1.x + f1.x, q1.y + f1.y, q1.z + f1.z), q3 = f3(q1.x + f2.x, q1.y + f2.y, q1.z + f2.z);
    F3 e0 = p2, e1 = p3 - p2, e2 = f3(-p3.x, -p3.y, -p3.z);
    F3 g0 = f1, g1 = f2 - f1, g2 = f3(-f2.x, -f2.y, -f2.z);
    const float Me = fmaxf(absMax(e0), fmaxf(absMax(e1), absMax(e2)));
    const float Mf = fmaxf(absMax(g0), fmaxf(absMax(g1), absMax(g2)));
    const float U = fmaxf(fmaxf(fmaxf(absMax(p2), absMax(p3)), fmaxf(absMax(q1), absMax(q2))), absMax(q3));
    const float S = fmaf(79.f * FLT_EPSILON, U, eta);
    const float eu = FLT_EPSILON * U;
    const float Kn = 120.f * eu * Me * Me, Km = 792.f * eu * Mf * Mf, Kef = 456.f * eu * Me * Mf;
    const float Kg = 384.f * eu * Me * Me * Me, Kh = 2400.f * eu * Mf * Mf * Mf;
    const F3 n1 = cross(e0, e1), m1 = cross(g0, g1);
    AxisState st{false, false};
    axisTest(n1, p2, p3, q1, q2, q3, S, Kn, st);
    axisTest(m1, p2, p3, q1, q2, q3, S, Km, st);
    // The other 15 axes as three rounds of five -- e_i x f_i, e_i x f_i+1, e_i x f_i+2, e_i x n1, f_i x m1 -- with the
    // edge vectors rotated through fixed registers between rounds: a third of the code of the unrolled form (the
    // kernels are bound by instruction fetch: their hot loop does not fit the 32 KB instruction cache), same axes,
    // same arithmetic per axis; the order of the axes does not matter to the verdict.
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        axisTest(cross(e0, g0), p2, p3, q1, q2, q3, S, Kef, st);
        axisTest(cross(e0, g1), p2, p3, q1, q2, q3, S, Kef, st);
        axisTest(cross(e0, g2), p2, p3, q1, q2, q3, S, Kef, st);
        axisTest(cross(e0, n1), p2, p3, q1, q2, q3, S, Kg, st);
        axisTest(cross(g0, m1), p2, p3, q1, q2, q3, S, Kh, st);
        const F3 te = e0, tg = g0;
        e0 = e1; e1 = e2; e2 = te;
        g0 = g1; g1 = g2; g2 = tg;
    }
    if (st.sep) return kTriNo;
    return st.unsure ? kTriUnsure : kTriYes;
}

// ---------------------------------------------------------------------------------------------


template <typename E>
struct Codec {
    int shB;  // entry = slot | a << 6 | b << shB
    __device__ __forceinline__ E pack(unsigned slot, unsigned a, unsigned b) const {
        return (E)slot | ((E)a << 6) | ((E)b << shB);
    }
    __device__ __forceinline__ unsigned slot(E e) const { return (unsigned)(e & 63); }
    __device__ __forceinline__ unsigned a(E e) const { return (unsigned)((e >> 6) & (((E)1 << (shB - 6)) - 1)); }
    __device__ __forceinline__ unsigned b(E e) const { return (unsigned)(e >> shB); }
};

struct Masks64 {
    unsigned lo, hi;
    __device__ __forceinline__ bool test(unsigned s) const { return ((s < 32 ? lo >> s : hi >> (s - 32)) & 1u) != 0; }
    __device__ __forceinline__ int count() const { return __popc(lo) + __popc(hi); }
};

// OR over the warp of (1 << slot) for the lanes with pred
__device__ __forceinline__ Masks64 slotBits(bool pred, unsigned slot) {
    Masks64 m;
    m.lo = __reduce_or_sync(kFull, (pred && slot < 32) ? 1u << slot : 0u);
    m.hi = __reduce_or_sync(kFull, (pred && slot >= 32) ? 1u << (slot - 32) : 0u);
    return m;
}

struct WarpCounters {
    unsigned long long bv = 0, tri = 0, exact = 0, states = 0;
    unsigned error = 0;
};

// ---- work sources ----------------------------------------------------------------------------
// A source hands out states in batches of up to 32 (warp-uniform control flow), names each with
// a (tag0, tag1) pair kept in the slot, can rebuild the FP64 state from the tags (exact path)
// and is told the verdict when the slot retires.

// isValid(q) over an array: states are taken in index order from a global counter.
template <bool kStreamed>
struct StateSourceT {
    static constexpr bool kIntervals = false;  // every slot is a single state
    const double *states;
    unsigned long long n;
    uint8_t *valid;
    unsigned long long *counter;
    // Host batches stream in while the kernel runs (`streamed`): the host copies the batch in pieces on one stream and,
    // behind each piece, writes the watermark word -- call epoch << 22 | states landed so far / 1024 -- with a stream
    // memory operation (cuStreamWriteValue32: ordered after the copy it follows, visible to running kernels).  States
    // are handed out in index order and a warp picks its claim up once the watermark covers the claim's last state
    // (acquire load at system scope).  The wait is bounded: a copy that never arrives is reported, not spun on.
    const double *hostStates = nullptr;   // zero-copy source (mapped page-locked host memory) or null
    const unsigned *water = nullptr;
    unsigned epoch = 0;
    unsigned long long landed = 0;   // states [0, landed) are known to be there
    static constexpr bool streamed = kStreamed;   // compile time: the resident kernel carries none of the streaming state
    unsigned stalled = 0;
    unsigned long long base = 0;
    bool exhausted = false;
    // Guided hand-out: a warp never takes more than its share of what is left (as of its previous
    // acquire), so the last states of a batch -- and all of a small batch -- spread over every warp of the
    // grid instead of queueing behind each other in a few (one warp tests ~32 box pairs per ~0.5 us, a
    // heavy pose needs ~1000 of them).
    unsigned long long crew = 1;   // guided divisor: warps in the grid x kGuidedShare
    unsigned long long seen = 0;   // counter value at this warp's previous acquire
#ifdef MPLB_WARP_TIMES
    unsigned long long dryAt = ~0ull;
#endif

    // streamed batches: a claim [pendBase, pendBase + pendCount) that has not landed yet, and the adaptive
    // claim size (halved when the warp had to wait for its data, doubled when the data was already there:
    // a kernel that outruns the copy keeps its claims just ahead of the copy engine instead of queueing
    // 64 states per warp behind it)
    unsigned long long pendBase = 0;
    int pendCount = 0, claim = 8;
    unsigned waitSpins = 0, backoff = 1000;   // back-off in SM cycles
    long long polledAt = 0;
    bool waited = false;

    __device__ bool done() const { return exhausted && pendCount == 0; }
    __device__ bool waiting() const { return pendCount > 0; }
    // mayBlock: the warp has nothing else to do (otherwise a claim that has not landed is simply retried later)
    template <typename SM>
    __device__ int acquire(int want, unsigned lane, SM &, bool mayBlock) {
        if (pendCount == 0) {
            if (exhausted) return 0;
            const unsigned long long left = n - min(seen, n);
            want = (int)min((unsigned long long)want, left / crew + 1);
            if (streamed && !hostStates) want = min(want, claim);
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(counter, (unsigned long long)want);
            b = __shfl_sync(kFull, b, 0);
            if (b >= n) {
                exhausted = true;
#ifdef MPLB_WARP_TIMES
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dryAt));
#endif
                return 0;
            }
            seen = b + want;
            if (n - b <= (unsigned long long)want) {
                exhausted = true;
                want = (int)(n - b);
            }
            if (!streamed) {
                base = b;
                return want;
            }
            if (hostStates) {
                // Zero-copy: the batch sits in page-locked host memory mapped into the device's address space.  The warp
                // pulls its claim over PCIe itself -- 56 B x want contiguous bytes, lane-contiguous 8-byte words, so the
                // requests are full lines -- and leaves a copy in device memory for the exact path's re-reads.  The
                // kernel's own progress paces the transfer; no copy engine, no arrival flags.
                const double *src = hostStates + 7 * b;
                double *dst = const_cast<double *>(states) + 7 * b;
                const int words = 7 * want;
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const int w = k * 32 + (int)lane;
                    if (w < words) __stcg(dst + w, __ldcs(src + w));
                }
                __syncwarp();   // the claim's device copy is complete for every lane of the warp
                base = b;
                return want;
            }
            pendBase = b;
            pendCount = want;
            waited = false;
        }
        // one lane looks at the watermark, at most once per back-off period so that thousands of waiting warps do not
        // crowd the L2 line it lives in
        const unsigned long long last = pendBase + pendCount - 1;
        int here = last < landed;
        if (!here) {
            if (waited && !mayBlock && clock64() - polledAt < (long long)backoff) return 0;
            // (a relaxed poll with a fence on every new watermark value was measured slower: 1.56 against 1.43 ms)
            unsigned w = 0;
            if (lane == 0) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(w) : "l"(water) : "memory");
            w = __shfl_sync(kFull, w, 0);
            __syncwarp();   // orders the other lanes' loads of the claim behind lane 0's acquire
            if ((w >> kWaterShift) == epoch) landed = (unsigned long long)(w & ((1u << kWaterShift) - 1u)) << kWaterUnitBits;
            here = last < landed;
        }
        if (!here) {
            waited = true;
            polledAt = clock64();
            if (mayBlock) __nanosleep(backoff / 2);   // ~2 cycles per ns
            backoff = min(backoff * 2, 16000u);
            if (++waitSpins > (1u << 14)) {   // ~0.13 s without a new piece (a piece takes 0.14 ms): reported, not spun on
                stalled = 1;
                pendCount = 0;
                exhausted = true;
            }
            return 0;
        }
        waitSpins = 0;
        backoff = 1000;
        claim = waited ? max(claim / 2, MPLB_CLAIM_MIN) : min(claim * 2, 32);
        base = pendBase;
        const int got = pendCount;
        pendCount = 0;
        return got;
    }
    struct Loader {
        const double *states;
        unsigned long long base;
        // -> false when the item turned out to need no work
        __device__ bool load(int rank, const uint4 &, double s[7], unsigned t[4], float &extraSlack) const {
            const unsigned long long i = base + rank;
            t[0] = (unsigned)i;
            t[1] = (unsigned)(i >> 32);
            t[2] = t[3] = 0;
            extraSlack = 0.f;
            Fetch{states}.stateOf(t[0], t[1], s);   // (streamed: the claim's watermark has been seen by acquire())
            return true;
        }
    };
    __device__ Loader loader() const { return Loader{states, base}; }
    struct Fetch {   // what it takes to read a state again (exact path), small enough to pass by value
        const double *states;
        __device__ void stateOf(unsigned t0, unsigned t1, double s[7]) const {
            const double *src = states + 7 * (((unsigned long long)t1 << 32) | t0);
            if (kStreamed) {   // freshly DMA-written memory: read through L2, not the non-coherent path
#pragma unroll
                for (int k = 0; k < 7; ++k) s[k] = __ldcg(src + k);
            } else {
#pragma unroll
                for (int k = 0; k < 7; ++k) s[k] = __ldg(src + k);
            }
        }
    };
    __device__ Fetch fetcher() const { return Fetch{states}; }
    __device__ void stateOf(unsigned t0, unsigned t1, double s[7]) const { fetcher().stateOf(t0, t1, s); }
    // warp-collective: slots `lane` (r0) and `lane + 32` (r1) retire
    template <typename SM>
    __device__ void retire(SM &sm, unsigned lane, bool r0, bool r1, const Masks64 &hit, const Masks64 &, unsigned &) const {
        if (r0) valid[((unsigned long long)sm.tag1[lane] << 32) | sm.tag0[lane]] = ((hit.lo >> lane) & 1u) ? 0 : 1;
        if (r1) valid[((unsigned long long)sm.tag1[lane + 32] << 32) | sm.tag0[lane + 32]] = ((hit.hi >> lane) & 1u) ? 0 : 1;
    }
    template <typename SM>
    __device__ void onHit(Masks64, SM &, Masks64 &, const Masks64 &) {}
};

// isValid(from,to) over an array of edges (se3...:134-178): verdict = valid(to) AND every interpolated
// sample i in [1, steps-1] at t = i * (1/steps).  The samples are not enumerated one by one.  Per edge there
// are 1 + 64 items: `to` itself, and the 64 equal INTERVALS [lo, hi] the samples [1, steps-1] are cut into
// (handed out in bit-reversed order, so the first ones spread over the edge and most invalid edges die
// early).  One traversal at an interval's middle sample, with every box test inflated by the largest
// displacement any robot point can have over the interval, either certifies all of its samples free at once
// (no leaf pair came within reach), or tests the middle sample exactly and hands the pieces left and right
// of it back (to the warp's interval LIFO, or to the next pass).  A sample is the middle of exactly one
// interval, so the worst case is one traversal per sample (as plain enumeration) and the free-space case is
// 65 traversals per edge.  The verdict is the reference's discrete one.
// Motion bound: translation is a lerp and Eigen's slerp turns at constant rate, so between samples i and m
// a robot point x moves at most |i-m|/steps * (||dt|| + 2 theta |x|), theta = acos|qa.qb|.
// pieces of the sample interval [a, b] (a <= b) after `extra` bisections into kids[cnt...], the split points last (they
// are popped, and can kill the edge, first).  Out of line and rolled up: it runs for a few lanes of a few retirements,
// and inlined four times with its loops unrolled it was a fifth of checkEdgesKernel's instructions -- a kernel that
// waits for instruction fetch more than for anything else.
__device__ __noinline__ unsigned splitHalf(unsigned e, unsigned a, unsigned b, int extra, uint4 *kids, unsigned cnt) {
    unsigned pa[4], pb[4], np = 1, pts[3], npts = 0;
    pa[0] = a; pb[0] = b;
#pragma unroll 1
    for (int lvl = 0; lvl < extra; ++lvl) {
        const unsigned cur = np;
#pragma unroll 1
        for (unsigned i = 0; i < cur; ++i) {
            const unsigned x = pa[i], y = pb[i];
            if (y - x < 2) continue;
            const unsigned m = x + (y - x) / 2;
            pts[npts++] = m;
            pb[i] = m - 1;       // m > x because y - x >= 2
            pa[np] = m + 1;      // m < y
            pb[np++] = y;
        }
    }
#pragma unroll 1
    for (unsigned i = 0; i < np; ++i) kids[cnt++] = make_uint4(e, pa[i], pb[i], 0);
#pragma unroll 1
    for (unsigned i = 0; i < npts; ++i) kids[cnt++] = make_uint4(e, pts[i], pts[i], 0);
    return cnt;
}
// (phase, edge) of item f when the item count needs 64 bits: the 64-bit division is a few hundred instructions
__device__ __noinline__ void itemOf64(unsigned long long f, unsigned long long nEdges, unsigned &ph, unsigned long long &e) {
    ph = (unsigned)(f / nEdges);
    e = f % nEdges;
}

struct EdgeSource {
    static constexpr bool kIntervals = true;   // slots may stand for a whole interval of samples
    const double *from, *to;
    const double *edgeConst;   // kEdgeConst doubles per edge (edgeConstKernel)
    const unsigned *steps;
    unsigned long long nEdges;
    unsigned long long crew = 1;   // warps in the grid x 2 (guided hand-out)
    unsigned *dead;
    unsigned long long *counter;
    uint4 *stack;          // this warp's interval LIFO (global memory): (edge, lo, hi, -)
    unsigned stackCap;
    float robotRadius;
    unsigned *errorFlag;
    // Passes 2 and 3 of a call take explicit (edge, lo, hi) descriptors: what the warps of the previous
    // pass spilled.  A warp that still holds pieces of a hard interval once the item pool has drained
    // hands the surplus on instead of grinding through it alone while the rest of the grid has gone home
    // (measured: one grazing 7905-step Cubicles edge kept two warps busy for 0.43 ms of a 0.48 ms launch).
    const uint4 *descIn = nullptr;
    unsigned long long nDesc = 0;
    uint4 *spillOut = nullptr;
    unsigned long long *spillCount = nullptr;
    unsigned spillCap = 0;
    bool globalDone = false;
    bool poolDry = false;   // the item counter has run past the last item (other warps may still hold work)
    unsigned stackCount = 0;
    unsigned long long first = 0;   // item counter as of this warp's previous grab (guided share)

    __device__ bool done() const { return globalDone && stackCount == 0; }

    // Up to `want` pieces of work as (edge, lo, hi) descriptors in sm.work[]: halves of intervals that could
    // not be certified first (depth first, keeps the LIFO short), topped up from as many items as it takes,
    // so that the FP64 state expansion runs on full warps.
    __device__ bool waiting() const { return false; }
    template <typename SM>
    __device__ int acquire(int want, unsigned lane, SM &sm, bool) {
        int count = min(want, (int)stackCount);
        stackCount -= count;
        if ((int)lane < count) sm.work[lane] = stack[stackCount + lane];
        // guided hand-out (as StateSource): new items are started only while the warp holds less than its
        // share of what is left, so that a small batch -- and the end of a large one -- spreads over the grid
        const unsigned long long nItems = descIn ? nDesc : (1 + kPhaseBSub) * nEdges;
        const int share = (int)min((unsigned long long)want, (nItems - min(first, nItems)) / crew + 1);
        if (!poolDry && count >= share) {   // busy with its own LIFO: still learn when the pool has drained
            unsigned long long c = 0;
            if (lane == 0) c = *(volatile unsigned long long *)counter;
            poolDry = __shfl_sync(kFull, c, 0) >= nItems;
        }
        if (descIn) {
            if (!globalDone && count < share) {
                const int take = min(want, share) - count;
                unsigned long long f = 0;
                if (lane == 0) f = atomicAdd(counter, (unsigned long long)take);
                f = __shfl_sync(kFull, f, 0);
                if (f >= nDesc) {
                    globalDone = poolDry = true;
                } else {
                    const int got = (int)min((unsigned long long)take, nDesc - f);
                    if ((int)lane < got) {
                        uint4 w = descIn[f + lane];
                        if (w.x >= nEdges) w = make_uint4(0, 1, 0, 0);   // never read out of range (lo > hi: no samples)
                        sm.work[count + lane] = w;
                    }
                    count += got;
                    first = f + got;
                }
            }
            __syncwarp();
            return count;
        }
        // implicit items: item f = (phase f / nEdges, edge f % nEdges), edge-minor so that neighbouring items are
        // different edges.  Each lane turns one item into its descriptor: one division per grab, none per item.
#pragma unroll 1
        for (int tries = 0; tries < 4 && count < share && !globalDone; ++tries) {
            const int take = min(want, share) - count;
            unsigned long long f = 0;
            if (lane == 0) f = atomicAdd(counter, (unsigned long long)take);
            f = __shfl_sync(kFull, f, 0);
            if (f >= nItems) {
                globalDone = poolDry = true;
                break;
            }
            first = f + take;
            unsigned long long e;
            unsigned ph;
            if ((nItems >> 32) == 0) {
                ph = (unsigned)f / (unsigned)nEdges;
                e = (unsigned)f - ph * (unsigned)nEdges;
            } else {
                itemOf64(f, nEdges, ph, e);
            }
            e += lane;
            while (e >= nEdges) {   // the grab runs into the next phase (lane < 32: more than once only for tiny batches)
                e -= nEdges;
                ++ph;
            }
            bool ok = false;
            unsigned lo = 0, hi = 0;
            if ((int)lane < take && f + lane < nItems && !((volatile unsigned *)dead)[e]) {
                const unsigned st = __ldg(steps + e);
                if (ph == 0) {   // sample `steps` stands for `to`
                    lo = hi = st;
                    ok = true;
                } else {
                    const unsigned long long m = st ? st - 1 : 0;   // samples 1 .. m
                    const unsigned k = __brev(ph - 1) >> (32 - kIntervalBits);   // bit reversal: 0, 32, 16, 48, ...
                    lo = 1u + (unsigned)((k * m) >> kIntervalBits);
                    hi = (unsigned)(((k + 1) * m) >> kIntervalBits);
                    ok = lo <= hi;
                }
            }
            const unsigned mk = __ballot_sync(kFull, ok);
            if (ok) sm.work[count + __popc(mk & ((1u << lane) - 1u))] = make_uint4((unsigned)e, lo, hi, 0);
            count += __popc(mk);
        }
        __syncwarp();
        return count;
    }
    // sample index `steps` stands for `to` itself
    struct Fetch {
        const double *from, *to, *edgeConst;
        const unsigned *steps;
        __device__ void stateOf(unsigned t0, unsigned t1, double s[7]) const {
            EdgeSource::sampleStateOf(from, to, edgeConst, t0, t1, __ldg(steps + t0), s);
        }
    };
    __device__ Fetch fetcher() const { return Fetch{from, to, edgeConst, steps}; }
    static __device__ void sampleStateOf(const double *from, const double *to, const double *edgeConst, unsigned e,
                                         unsigned i, unsigned st, double s[7]) {
        const double *b = to + 7 * (size_t)e;
        if (i >= st) {
#pragma unroll
            for (int k = 0; k < 7; ++k) s[k] = __ldg(b + k);
            return;
        }
        const double *a = from + 7 * (size_t)e, *c = edgeConst + kEdgeConst * (size_t)e;
        double fa[7], fb[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            fa[k] = __ldg(a + k);
            fb[k] = __ldg(b + k);
        }
        interpolateWith(fa, fb, dmul((double)i, __ddiv_rn(1.0, (double)st)), __ldg(c), __ldg(c + 1), __ldg(c + 2), s);
    }
    struct Loader {
        const double *from, *to, *edgeConst;
        const unsigned *steps, *dead;
        __device__ bool load(int, const uint4 &w, double s[7], unsigned t[4], float &extraSlack) const {
            const unsigned e = w.x, lo = w.y, hi = w.z;
            if (((volatile const unsigned *)dead)[e]) return false;
            const unsigned st = __ldg(steps + e);
            const unsigned mid = lo + (hi - lo) / 2;
            EdgeSource::sampleStateOf(from, to, edgeConst, e, mid, st, s);
            extraSlack = 0.f;
            if (hi > lo) {
                const double *c = edgeConst + kEdgeConst * (size_t)e;
                const double reach = (double)max(hi - mid, mid - lo) / (double)st;
                extraSlack = __double2float_ru((reach * __ldg(c + 3)) * 1.000001 + __ldg(c + 4));
            }
            t[0] = e; t[1] = mid; t[2] = lo; t[3] = hi;
            return true;
        }
    };
    __device__ Loader loader() const { return Loader{from, to, edgeConst, steps, dead}; }

    // warp-collective.  A hit kills the edge; an interval whose traversal reached leaf pairs (touched) and
    // whose middle sample is free hands its two halves to the LIFO; anything else is finished.
    template <typename SM>
    __device__ void retire(SM &sm, unsigned lane, bool r0, bool r1, const Masks64 &hit, const Masks64 &touched,
                           unsigned &error) {
        unsigned cnt = 0;
        uint4 kids[28];
        // With the item pool drained and little left on the LIFO the warp is latency bound (a hard interval
        // is a chain of ~log2(length) traversals, ~50 us each): split each half `extra` more times then --
        // the inner split points are tested as single samples next to the pieces, which skips the attempts
        // to certify the larger pieces and shortens the chain (measured: 256 Cubicles edges 0.63 -> 0.50 ms
        // with one extra level).
        const int extra = !(poolDry || globalDone) ? 0 : stackCount < 8 ? 2 : stackCount < 32 ? 1 : 0;
        auto one = [&](bool r, unsigned slot, bool isHit, bool isTouched) {
            if (!r) return;
            const unsigned e = sm.tag0[slot], mid = sm.tag1[slot], lo = sm.tag2[slot], hi = sm.tag3[slot];
            if (isHit) {
                dead[e] = 1u;
                return;
            }
            if (hi > lo && isTouched && !((volatile unsigned *)dead)[e]) {
                if (mid > lo) cnt = splitHalf(e, lo, mid - 1, extra, kids, cnt);
                if (mid < hi) cnt = splitHalf(e, mid + 1, hi, extra, kids, cnt);
            }
        };
#pragma unroll 1
        for (int h = 0; h < 2; ++h)   // (rolled up: one copy of the slot's code, see splitHalf)
            one(h ? r1 : r0, lane + 32u * h, ((h ? hit.hi : hit.lo) >> lane) & 1u, ((h ? touched.hi : touched.lo) >> lane) & 1u);
        // exclusive prefix of cnt over the warp
        unsigned incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(kFull, incl, d);
            if ((int)lane >= d) incl += v;
        }
        const unsigned total = __shfl_sync(kFull, incl, 31);
        if (total == 0) return;
        if (spillOut && (poolDry || globalDone) && stackCount + total > kSpillKeep) {
            // One atomicAdd per reservation and no reservation is ever undone (a compare-and-swap loop on the one
            // counter made thousands of warps retry against each other: 65,536 Cubicles edges 1.0 -> 3.3 ms).  The
            // count may run past the capacity; the next pass reads min(count, cap) descriptors.  At most one failed
            // reservation starts below the capacity: that warp fills [pos, cap) with empty descriptors (lo > hi), so
            // every slot the next pass reads was written by this pass, and keeps its pieces on its own LIFO.
            unsigned long long pos = 0;
            if (lane == 0) pos = atomicAdd(spillCount, (unsigned long long)total);
            pos = __shfl_sync(kFull, pos, 0);
            if (pos + total <= spillCap) {
#pragma unroll 1
                for (unsigned k = 0; k < cnt; ++k) spillOut[pos + incl - cnt + k] = kids[k];
                return;
            }
            for (unsigned long long at = pos + lane; at < spillCap; at += 32) spillOut[at] = make_uint4(0, 1, 0, 0);
        }
        if (stackCount + total > stackCap) {
            error = 1;  // reported by the host call; never silent
            return;
        }
        const unsigned at = stackCount + incl - cnt;
#pragma unroll 1
        for (unsigned k = 0; k < cnt; ++k) stack[at + k] = kids[k];
        stackCount += total;
        __syncwarp();
    }
    // a sample of an edge collided: drop the rest of that edge held by this warp
    template <typename SM>
    __device__ void onHit(Masks64 fresh, SM &sm, Masks64 &hit, const Masks64 &inUse) {
        const unsigned lane = threadIdx.x & 31u;
        while (fresh.lo | fresh.hi) {
            unsigned s;
            if (fresh.lo) {
                s = __ffs(fresh.lo) - 1;
                fresh.lo &= fresh.lo - 1;
            } else {
                s = 32 + __ffs(fresh.hi) - 1;
                fresh.hi &= fresh.hi - 1;
            }
            const unsigned e = sm.tag0[s];
            if (lane == 0) dead[e] = 1u;
            const bool k0 = inUse.test(lane) && sm.tag0[lane] == e;
            const bool k1 = kSlots > 32 && inUse.test(lane + 32) && sm.tag0[lane + 32] == e;
            hit.lo |= __ballot_sync(kFull, k0);
            hit.hi |= __ballot_sync(kFull, k1);
        }
    }
};

// ---- the per-warp engine -----------------------------------------------------------------------

// kLatency: the variant launched for small pose batches, where the time to the last verdict is the latency of the
// heaviest pose and not throughput (see the end of the step loop)
template <typename E, bool kLatency = false, bool kMargins = false, typename Source>
__device__ void runWarp(WarpSmem<E> &sm, const SceneDev &sc, Source &src, WarpCounters &cnt) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned ltMask = (1u << lane) - 1u;
    const Codec<E> cd{6 + sc.bitsA};
    const int bitsA = sc.bitsA;
    const unsigned maskA = (1u << bitsA) - 1u;
    Masks64 freeM{~0u, kSlots > 32 ? ~0u : 0u}, readyM{0, 0}, hitM{0, 0}, doneM{0, 0}, touchedM{0, 0};  // touched: reached a leaf pair
    // this lane's LIFO holds entries [bot, sp); thieves take from the bottom (the largest subtree)
    int nt = 0, nx = 0, sp = 0, bot = 0, mySlot = -1;
    const int triMin = sc.triBatchMin, triEager = sc.triEagerLanes;
    unsigned idleSpins = 0, stepsLane = 0;
#ifndef MPLB_FIXED_REUSE
#define MPLB_FIXED_REUSE 0
#endif
    // The fixed box of a lane's step is, more often than not, the fixed box of its previous step: the nearer surviving
    // child is popped right away, and when it is that child's side that gets split again the other box stays (counted
    // with -DMPLB_COUNT_FIXED: 52 % of the steps on Alpha 1.5, 78 % on Twistycool, 84 % on Cubicles).  With
    // -DMPLB_FIXED_REUSE=1 its frame (ObbFrame: 15 floats), link and size are kept and such a step skips two of its six
    // node loads and the slot's transform rows.  Built, verified against the oracle and measured: 2^20 poses Alpha
    // 0.738 -> 0.744 ms, Cubicles 0.950 -> 0.931, Apartment 3.20 -> 3.15; edge batches 1 % slower.  A fifth fewer L1
    // wavefronts buy nothing: the warps wait on each step's dependent chain (load -> frame -> 15 axes -> votes), not on
    // the L1 pipe -- hence off.
    ObbFrame keptF;
    unsigned keptKey = ~0u;   // fixed node | role << 31 the kept frame belongs to (~0: none; reset when the slot changes)
    int keptLink = 0;
    float keptSz = 0.f;
#ifdef MPLB_COUNT_FIXED
    unsigned dbgPrev = ~0u, dbgSame = 0;
    int dbgPrevSlot = -1;
#endif
    // compact the set bits of a 64-slot mask into sm.list (lane s looks after slots s and s + 32)
    auto buildList = [&](const Masks64 &m) {
        if ((m.lo >> lane) & 1u) sm.list[__popc(m.lo & ltMask)] = lane;
        if ((m.hi >> lane) & 1u) sm.list[__popc(m.lo) + __popc(m.hi & ltMask)] = lane + 32;
        __syncwarp();
    };

    for (;;) {
        // ---- a lane whose LIFO ran dry leaves its slot; the last one out ends the slot's search
        {
            const bool fin = mySlot >= 0 && sp == bot;
            bool last = false;
            if (fin) last = atomicSub(&sm.workers[mySlot], 1) == 1;
            const Masks64 d = slotBits(last, (unsigned)mySlot);
            doneM.lo |= d.lo;
            doneM.hi |= d.hi;
            if (fin) {
                mySlot = -1;
                sp = bot = 0;
            }
        }
        // ---- retire finished slots whose queued leaf pairs are all resolved
        {
            const bool r0 = ((doneM.lo >> lane) & 1u) && sm.pending[lane] == 0;
            const bool r1 = kSlots > 32 && ((doneM.hi >> lane) & 1u) && sm.pending[lane + 32] == 0;
            const unsigned m0 = __ballot_sync(kFull, r0), m1 = __ballot_sync(kFull, r1);
            if (m0 | m1) {
                src.retire(sm, lane, r0, r1, hitM, touchedM, cnt.error);
                freeM.lo |= m0; doneM.lo &= ~m0; hitM.lo &= ~m0; touchedM.lo &= ~m0;
                freeM.hi |= m1; doneM.hi &= ~m1; hitM.hi &= ~m1; touchedM.hi &= ~m1;
            }
        }
        int nActive = __popc(__ballot_sync(kFull, sp > bot));
        unsigned needMask = __ballot_sync(kFull, mySlot < 0);
        // ---- expand a batch of new states into free slots (all lanes convert in parallel)
        if (!src.done() && needMask && !(readyM.lo | readyM.hi)) {
            const int nfree = freeM.count();
            if (nfree >= 16 || (nfree > 0 && nActive <= 16)) {
                const int got = src.acquire(min(32, nfree), lane, sm, nActive == 0 && nt == 0 && nx == 0);
                buildList(freeM);
                bool take = (int)lane < got;
                unsigned slot = 0;
                if (take) {
                    slot = sm.list[lane];
                    take = fillSlot<Source::kIntervals && MPLB_EDGE_FILL_OUTLINE>(src.loader(), Source::kIntervals ? sm.work[lane] : make_uint4(0, 0, 0, 0), (int)lane, sc.envRadius, sc.robotRadius, sm.xf[slot],
                                    sm.tag0 + slot);
                }
                const Masks64 t = slotBits(take, slot);
                freeM.lo &= ~t.lo; freeM.hi &= ~t.hi;
                readyM.lo |= t.lo; readyM.hi |= t.hi;
                cnt.states += __popc(__ballot_sync(kFull, take));
                __syncwarp();
            }
        }
        // ---- idle lanes adopt ready slots
        if (needMask && (readyM.lo | readyM.hi)) {
            buildList(readyM);
            const int rank = __popc(needMask & ltMask);
            const bool take = mySlot < 0 && rank < readyM.count();
            unsigned slot = 0;
            if (take) {
                slot = sm.list[rank];
                mySlot = (int)slot;
                keptKey = ~0u;
                sm.workers[slot] = 1;
                if (sc.haveGeometry && !sc.bothRootsLeaf) {
                    sm.stack[0][lane] = sc.rootEntry;  // root pair, expanded without testing itself
                    bot = 0;
                    sp = 1;
                }
            }
            if (sc.bothRootsLeaf) {  // two single-triangle meshes: straight to the leaf test
                const unsigned m = __ballot_sync(kFull, take);
                if (take) {
                    sm.triq[nt + __popc(m & ltMask)] = cd.pack(slot, 0, 0);
                    sm.pending[slot] = 1;
                }
                nt += __popc(m);
                const Masks64 tm = slotBits(take, slot);
                touchedM.lo |= tm.lo; touchedM.hi |= tm.hi;
            }
            const Masks64 t = slotBits(take, slot);
            readyM.lo &= ~t.lo; readyM.hi &= ~t.hi;
            __syncwarp();
            needMask = __ballot_sync(kFull, mySlot < 0);
        }
        // ---- work stealing: lanes still idle take the bottom pair (the largest pending subtree) of
        //      lanes that have at least two, and join that slot's search
        if (needMask) {
            const bool canGive = sp - bot >= 2;
            const unsigned donors = __ballot_sync(kFull, canGive);
            const int k = min(__popc(needMask), __popc(donors));
            if (k > 0) {
                const int donorRank = __popc(donors & ltMask);
                const bool give = canGive && donorRank < k;
                if (give) sm.list[donorRank] = lane;
                __syncwarp();
                const int idleRank = __popc(needMask & ltMask);
                const bool steal = mySlot < 0 && idleRank < k;
                const unsigned dl = steal ? sm.list[idleRank] : lane;
                const int dbot = __shfl_sync(kFull, bot, dl), dslot = __shfl_sync(kFull, mySlot, dl);
                if (steal) {
                    sm.stack[0][lane] = sm.stack[dbot][dl];
                    bot = 0;
                    sp = 1;
                    mySlot = dslot;
                    keptKey = ~0u;
                    atomicAdd(&sm.workers[dslot], 1);
                }
                if (give) ++bot;
                __syncwarp();
            }
        }
        nActive = __popc(__ballot_sync(kFull, sp > bot));
        if (nActive == 0 && nt == 0 && nx == 0) {
            const bool busy = __any_sync(kFull, mySlot >= 0) || (doneM.lo | doneM.hi) || (readyM.lo | readyM.hi);
            if (!busy && src.done()) break;
            if (src.waiting()) idleSpins = 0;   // a streamed claim that has not landed: the source bounds that wait itself
            if (++idleSpins > 4096u) {  // bookkeeping bug guard: report instead of hanging the device
                cnt.error = 1;
                break;
            }
            continue;
        }
        idleSpins = 0;

        // ---- triangle-triangle batch (FP32 filter)
        auto triBatch = [&]() {
            const int n = min(32, nt);
            const bool act0 = (int)lane < n;
            const E en = act0 ? sm.triq[nt - 1 - lane] : (E)0;
            nt -= n;
            __syncwarp();
            const unsigned slot = cd.slot(en);
            const bool act = act0 && !hitM.test(slot);
            int verdict = kTriNo;
            if (act) {
                const float4 *tp = sc.robotTris + 4 * (size_t)cd.a(en);
                const float4 *tq = sc.envTris + 4 * (size_t)cd.b(en);
                float4 p1, p2, w1, w2;
                ldg8(tp, p1, p2);
                ldg8(tq, w1, w2);
                const float4 p3 = ldg4(tp + 2), w3 = ldg4(tq + 2);
                const float4 x0 = sm.xf[slot][0], x1 = sm.xf[slot][1], x2 = sm.xf[slot][2], x3 = sm.xf[slot][3];
                // q1 = R*Q1 + T - P1 ; env edge vectors are only rotated
                const F3 q1 = f3(fmaf(x0.x, w1.x, fmaf(x0.y, w1.y, fmaf(x0.z, w1.z, x0.w))) - p1.x,
                                 fmaf(x1.x, w1.x, fmaf(x1.y, w1.y, fmaf(x1.z, w1.z, x1.w))) - p1.y,
                                 fmaf(x2.x, w1.x, fmaf(x2.y, w1.y, fmaf(x2.z, w1.z, x2.w))) - p1.z);
                auto rot = [&](const float4 w) {
                    return f3(fmaf(x0.x, w.x, fmaf(x0.y, w.y, x0.z * w.z)), fmaf(x1.x, w.x, fmaf(x1.y, w.y, x1.z * w.z)),
                              fmaf(x2.x, w.x, fmaf(x2.y, w.y, x2.z * w.z)));
                };
                verdict = triTriFilter(f3(p2.x, p2.y, p2.z), f3(p3.x, p3.y, p3.z), q1, rot(w2), rot(w3), x3.y);
            }
            cnt.tri += __popc(__ballot_sync(kFull, act));
            if (act0 && verdict != kTriUnsure) atomicSub(&sm.pending[slot], 1);
            const Masks64 fresh = slotBits(verdict == kTriYes, slot);
            const unsigned mu = __ballot_sync(kFull, verdict == kTriUnsure);
            if (mu) {
                if (verdict == kTriUnsure) sm.exq[nx + __popc(mu & ltMask)] = en;
                nx += __popc(mu);
            }
            __syncwarp();
            if (fresh.lo | fresh.hi) {
                hitM.lo |= fresh.lo;
                hitM.hi |= fresh.hi;
                src.onHit(fresh, sm, hitM, Masks64{~freeM.lo, ~freeM.hi});
            }
        };
        // ---- exact FP64 batch
        auto exactBatch = [&]() {
            const int n = min(32, nx);
            const bool act0 = (int)lane < n;
            const E en = act0 ? sm.exq[nx - 1 - lane] : (E)0;
            nx -= n;
            __syncwarp();
            const unsigned slot = cd.slot(en);
            const bool act = act0 && !hitM.test(slot);
            bool hit = false;
            if (act)
                // (the margin-recording twin is only ever referenced from the kMargins instantiations: its mere presence in
                // the call graph of the regular kernels cost them 25 % -- 0.80 -> 1.01 ms per 2^20 Alpha poses)
                if constexpr (kMargins)
                    hit = exactPairHitMargin(src.fetcher(), sm.tag0[slot], sm.tag1[slot], sc.robotTris64 + 9 * (size_t)cd.a(en),
                                             sc.envTris64 + 9 * (size_t)cd.b(en), sc.minMargin);
                else
                    hit = exactPairHit(src.fetcher(), sm.tag0[slot], sm.tag1[slot], sc.robotTris64 + 9 * (size_t)cd.a(en),
                                       sc.envTris64 + 9 * (size_t)cd.b(en));
            cnt.exact += __popc(__ballot_sync(kFull, act));
            if (act0) atomicSub(&sm.pending[slot], 1);
            const Masks64 fresh = slotBits(hit, slot);
            __syncwarp();
            if (fresh.lo | fresh.hi) {
                hitM.lo |= fresh.lo;
                hitM.hi |= fresh.hi;
                src.onHit(fresh, sm, hitM, Masks64{~freeM.lo, ~freeM.hi});
            }
        };
        // The queues hold at most 31 leftovers plus what one step can add (64 leaf pairs, 32 unsure
        // pairs per triangle batch): drain them down below a batch before stepping again.
        if constexpr (Source::kIntervals) {
            // (edges: one call site per batch kind -- they are a few hundred instructions apiece and that kernel's loop
            // does not fit the instruction cache: 65,536 Cubicles edges 0.790 -> 0.761 ms; poses measured 1 % slower)
            for (;;) {
                const bool wantTri = nt >= triMin || (nt > 0 && nActive <= triEager);
                if (nx >= 32 || (!wantTri && nx > 0 && nActive <= triEager)) exactBatch();
                else if (wantTri) triBatch();
                else break;
            }
        } else {
            while (nt >= triMin || (nt > 0 && nActive <= triEager)) {
                triBatch();
                while (nx >= 32) exactBatch();
            }
            if (nx > 0 && nActive <= triEager) exactBatch();
        }
        // ---- depth-first steps: each lane pops a pair and tests both children of the larger box.
        //      A few steps run back to back before the slot bookkeeping above is revisited (a lane that
        //      finishes early idles for at most kStepsPerRound - 1 steps; the leaf queue is re-checked
        //      every step because one step can add 64 pairs to it).  Stealing more often than that does
        //      not pay: it fans a colliding pose's search out over subtrees the nearest-first descent
        //      would never have visited (measured: +65 % box tests on Alpha 1.5).
        for (int rep = 0; rep < kStepsPerRound && nActive && nt < 32; ++rep) {
            bool act = sp > bot;
            const unsigned slot = (unsigned)max(mySlot, 0);
            if (act && hitM.test(slot)) {  // decided meanwhile (leaf test, another lane, or a sibling sample of the edge)
                sp = bot;
                act = false;
            }
            bool ov0 = false, ov1 = false, plain0 = false, plain1 = false, leaf0 = false, leaf1 = false, descendA = false;
            float d0 = 0.f, d1 = 0.f, sz0 = 0.f, sz1 = 0.f, szF = 0.f;
            unsigned fixedIdx = 0, child = 0;
            int link0 = 0, link1 = 0, fixedLink = 0;
            if (act) {
                // entry: robotIdx | envIdx << bitsA | descendRobot << 31, where the index on the
                // descended side already is that node's first child: every load below is
                // independent of the others (one memory round trip per step)
                const unsigned en = sm.stack[--sp][lane];
                descendA = (en >> 31) != 0;
                const unsigned r = en & maskA, e = (en & 0x7fffffffu) >> bitsA;
                child = descendA ? r : e;
                fixedIdx = descendA ? e : r;
#ifdef MPLB_COUNT_FIXED   // diagnostic: how often a lane's next step has the same fixed box as its previous one
                if (dbgPrev == (fixedIdx | (descendA ? 0x80000000u : 0u)) && dbgPrevSlot == (int)slot) ++dbgSame;
                dbgPrev = fixedIdx | (descendA ? 0x80000000u : 0u);
                dbgPrevSlot = (int)slot;
#endif
                const float4 *pc = (descendA ? sc.robotNodes : sc.envNodes) + 4 * (size_t)child;
                const float4 *pf = (descendA ? sc.envNodes : sc.robotNodes) + 4 * (size_t)fixedIdx;
                float4 c00, c01, c02, c0e, c10, c11, c12, c1e;
                const unsigned key = fixedIdx | (descendA ? 0x80000000u : 0u);
                const bool kept = MPLB_FIXED_REUSE && key == keptKey;
                ldg8(pc, c00, c01); ldg8(pc + 2, c02, c0e);
                ldg8(pc + 4, c10, c11); ldg8(pc + 6, c12, c1e);
                const float4 x3 = sm.xf[slot][3];
                if (!kept) {
                    float4 f0, f1, f2, fe;
                    ldg8(pf, f0, f1); ldg8(pf + 2, f2, fe);
                    const float4 x0 = sm.xf[slot][0], x1 = sm.xf[slot][1], x2 = sm.xf[slot][2];
                    // The fixed box is always the first argument.  (R|T) maps env -> robot coordinates; when
                    // the fixed box is the environment's the test runs in its frame with the inverse map
                    // (R^T | -R^T T), so no per-role shuffling of the 3 x 16 node floats is needed.
                    const float4 y0 = descendA ? make_float4(x0.x, x1.x, x2.x, -fmaf(x0.x, x0.w, fmaf(x1.x, x1.w, x2.x * x2.w))) : x0;
                    const float4 y1 = descendA ? make_float4(x0.y, x1.y, x2.y, -fmaf(x0.y, x0.w, fmaf(x1.y, x1.w, x2.y * x2.w))) : x1;
                    const float4 y2 = descendA ? make_float4(x0.z, x1.z, x2.z, -fmaf(x0.z, x0.w, fmaf(x1.z, x1.w, x2.z * x2.w))) : x2;
                    // what depends on the fixed box and the pose only is formed once for both children (device_math.cuh)
                    obbFrame(f0, f1, f2, fe, y0, y1, y2, keptF);
                    keptLink = __float_as_int(fe.w);
                    keptSz = fmaf(fe.x, fe.x, fmaf(fe.y, fe.y, fe.z * fe.z));
                    keptKey = key;
                }
                const ObbFrame &F = keptF;
                float g0 = 0.f, g1 = 0.f;
#ifndef MPLB_OBB_LOOP
#define MPLB_OBB_LOOP 0
#endif
                if constexpr (Source::kIntervals || MPLB_OBB_LOOP) {
                    // one copy of the per-child test, run twice: checkEdgesKernel is bound by instruction fetch (ncu:
                    // stall_no_inst; its warps alternate between this loop and the FP64 sample conversion), not by
                    // the dependent FMA chains the two inlined copies overlap
                    float4 b0 = c00, b1 = c01, b2 = c02, be = c0e;
#pragma unroll 1
                    for (int k = 0; k < 2; ++k) {
                        float dd;
                        const float g = obbGapFramed(F, b0, b1, b2, be, dd);
                        if (k == 0) {
                            g0 = g; d0 = dd;
                            b0 = c10; b1 = c11; b2 = c12; be = c1e;
                        } else {
                            g1 = g; d1 = dd;
                        }
                    }
                } else {
                    g0 = obbGapFramed(F, c00, c01, c02, c0e, d0);
                    g1 = obbGapFramed(F, c10, c11, c12, c1e, d1);
                }
                ov0 = g0 <= x3.x;   // within reach (plain slack, or inflated for an interval certificate)
                ov1 = g1 <= x3.x;
                if constexpr (Source::kIntervals) {
                    plain0 = g0 <= x3.z;  // overlapping at this very sample
                    plain1 = g1 <= x3.z;
                } else {
                    plain0 = ov0;
                    plain1 = ov1;
                }
                link0 = __float_as_int(c0e.w);
                link1 = __float_as_int(c1e.w);
                fixedLink = keptLink;
                const bool fixedLeaf = fixedLink < 0;
                leaf0 = fixedLeaf && link0 < 0;
                leaf1 = fixedLeaf && link1 < 0;
                sz0 = fmaf(c0e.x, c0e.x, fmaf(c0e.y, c0e.y, c0e.z * c0e.z));
                sz1 = fmaf(c1e.x, c1e.x, fmaf(c1e.y, c1e.y, c1e.z * c1e.z));
                szF = keptSz;
            }
            stepsLane += act ? 1u : 0u;   // box tests are counted per lane and summed when the warp is through
            // leaf-leaf survivors -> shared triangle queue
            // a leaf pair within reach means the interval cannot be certified (touched); only pairs that
            // overlap at this very sample need the triangle test
            const bool reach0 = ov0 && leaf0, reach1 = ov1 && leaf1;
            const bool t0 = reach0 && plain0, t1 = reach1 && plain1;
            const unsigned m0 = __ballot_sync(kFull, t0), m1 = __ballot_sync(kFull, t1);
            if (m0 | m1) {
                const unsigned fixedTri = (unsigned)(-(fixedLink + 1));
                if (t0) {
                    const unsigned ct = (unsigned)(-(link0 + 1));
                    sm.triq[nt + __popc(m0 & ltMask)] = descendA ? cd.pack(slot, ct, fixedTri) : cd.pack(slot, fixedTri, ct);
                }
                if (t1) {
                    const unsigned ct = (unsigned)(-(link1 + 1));
                    sm.triq[nt + __popc(m0) + __popc(m1 & ltMask)] = descendA ? cd.pack(slot, ct, fixedTri) : cd.pack(slot, fixedTri, ct);
                }
                if (t0 || t1) atomicAdd(&sm.pending[slot], (t0 ? 1 : 0) + (t1 ? 1 : 0));
                nt += __popc(m0) + __popc(m1);
                if (nt > kQueueCap) cnt.error = 1;  // cannot happen (see the drain loop above); never silent
            }
            if (Source::kIntervals && __any_sync(kFull, reach0 || reach1)) {
                // the rest of a touched interval's search only has to decide its middle sample: drop the inflation
                if (reach0 || reach1)
                    reinterpret_cast<float *>(&sm.xf[slot][3])[0] = reinterpret_cast<const float *>(&sm.xf[slot][3])[2];
                const Masks64 tm = slotBits(reach0 || reach1, slot);
                touchedM.lo |= tm.lo; touchedM.hi |= tm.hi;
            }
            // inner survivors -> own LIFO, farther one first so the nearer one is searched first.
            // The pushed entry already names the first child of whichever box the next step splits:
            // the larger one unless it is a leaf ([EXT] FCL's firstOverSecond rule).
            const bool s0 = ov0 && !leaf0, s1 = ov1 && !leaf1;
            if (s0 || s1) {
                auto entryFor = [&](unsigned childIdx, int childLink, float childSz) {
                    // robot side / env side of the surviving pair
                    const int linkR = descendA ? childLink : fixedLink, linkE = descendA ? fixedLink : childLink;
                    const unsigned idxR = descendA ? childIdx : fixedIdx, idxE = descendA ? fixedIdx : childIdx;
                    const float szR = descendA ? childSz : szF, szE = descendA ? szF : childSz;
                    const bool nextA = linkE < 0 || (linkR >= 0 && szR > szE);
                    return nextA ? ((unsigned)linkR | idxE << bitsA | 0x80000000u) : (idxR | (unsigned)linkE << bitsA);
                };
                const unsigned e0 = entryFor(child, link0, sz0), e1 = entryFor(child + 1, link1, sz1);
                unsigned top;
                if (s0 && s1) {
                    const bool zeroFirst = d0 > d1;  // push the farther first
                    sm.stack[sp][lane] = zeroFirst ? e0 : e1;
                    sm.stack[sp + 1][lane] = top = zeroFirst ? e1 : e0;
                    sp += 2;
                } else {
                    sm.stack[sp][lane] = top = s0 ? e0 : e1;
                    sp += 1;
                }
                // the pair this lane pops next: start pulling its three nodes towards L1 now (children
                // pairs are 128-byte aligned, so one line holds both)
                {
                    const bool dA = (top >> 31) != 0;
                    const unsigned r = top & maskA, e = (top & 0x7fffffffu) >> bitsA;
                    const float4 *pc = dA ? sc.robotNodes + 4 * (size_t)r : sc.envNodes + 4 * (size_t)e;
                    const float4 *pf = dA ? sc.envNodes + 4 * (size_t)e : sc.robotNodes + 4 * (size_t)r;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pc));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
                }
            }
            __syncwarp();
            nActive = __popc(__ballot_sync(kFull, sp > bot));
            // Once the batch is handed out nothing but stealing can fill idle lanes, and the last poses' latency is
            // all that is left: go back to the bookkeeping as soon as an idle lane could take a pair from a busy one
            // (1 024 random Alpha poses 75 -> 57 us, a 1 000-test pose alone 55 -> 45 us).  Not in the throughput
            // variant: the two extra votes per step cost 3 % on 2^20-pose batches.
#ifndef MPLB_TAIL_STEAL
#define MPLB_TAIL_STEAL 0
#endif
            if constexpr (kLatency || MPLB_TAIL_STEAL)
                if (src.done() && nActive < 32 && __any_sync(kFull, sp - bot >= 2)) break;
        }
    }
    cnt.bv += 2ull * __reduce_add_sync(kFull, stepsLane);
#ifdef MPLB_COUNT_FIXED
    cnt.exact += __reduce_add_sync(kFull, dbgSame);   // (reported in place of the exact-test count)
#endif
}

template <typename E>
__device__ __forceinline__ WarpSmem<E> warpSmem(int levels) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    WarpSmem<E> sm;
    sm.carve(smemRaw + (threadIdx.x >> 5) * WarpSmem<E>::bytes(levels), levels);
    sm.pending[threadIdx.x & 31u] = 0;
    if (kSlots > 32) sm.pending[(threadIdx.x & 31u) + 32] = 0;
    __syncwarp();
    return sm;
}

__device__ __forceinline__ void flushCounters(const WarpCounters &c, DevCounters *ctr, bool countStates) {
    if ((threadIdx.x & 31u) == 0) {
        atomicAdd(&ctr->bvTests, c.bv);
        atomicAdd(&ctr->triTests, c.tri);
        atomicAdd(&ctr->exactTests, c.exact);
        if (countStates) atomicAdd(&ctr->states, c.states);
        if (c.error) atomicExch(&ctr->overflow, 1u);
    }
}

// Direct-return calls (HostReturn): every warp publishes its verdict stores and counter updates, the grid's last warp
// hands the counters to the host, leaves the device control block clean for the next call and rings the host.
__device__ __forceinline__ void ringHost(const HostReturn &ret, DevCounters *ctr) {
    if (!ret.flag) return;
    const unsigned lane = threadIdx.x & 31u;
    __threadfence_system();
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(&ctr->ticket, 1u);
    t = __shfl_sync(kFull, t, 0);
    if (t != gridDim.x * (blockDim.x >> 5) - 1) return;
    __threadfence();
    const unsigned words = sizeof(DevCounters) / 4;
    unsigned *src = reinterpret_cast<unsigned *>(ctr), *dst = reinterpret_cast<unsigned *>(ret.hostCtr);
    for (unsigned w = lane; w < words; w += 32) {
        dst[w] = __ldcg(src + w);
        src[w] = 0u;
    }
    for (unsigned w = lane; w < ret.tailWords; w += 32) dst[ret.tailAt + w] = __ldcg(src + ret.tailAt + w);
    for (unsigned w = lane; w < ret.clearWords; w += 32) src[words + w] = 0u;
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
        const unsigned seq = ret.seqAt ? *(const volatile unsigned *)ret.seqAt : ret.seq;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ret.flag), "r"(seq) : "memory");
    }
}

template <typename E, bool kLatency, bool kStreamed, bool kMargins = false>
__global__ void __launch_bounds__(kWarpsStates * 32, MPLB_MIN_CTAS)
checkStatesKernel(SceneDev sc, const double *states, unsigned long long n, uint8_t *__restrict__ valid,
                  DevCounters *ctr, unsigned long long *workCounter, unsigned guidedShare, StreamedInput in) {
    WarpSmem<E> sm = warpSmem<E>(sc.stackLevels);
    StateSourceT<kStreamed> src;
    src.states = states; src.n = n; src.valid = valid; src.counter = workCounter;
    src.hostStates = in.host; src.water = in.water; src.epoch = in.epoch;
    src.crew = (unsigned long long)gridDim.x * (blockDim.x >> 5) * guidedShare;
    WarpCounters cnt;
#ifdef MPLB_WARP_TIMES
    auto now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if ((threadIdx.x & 31u) == 0) atomicMax(&ctr->negStart, ~now());
#endif
    runWarp<E, kLatency, kMargins>(sm, sc, src, cnt);
#ifdef MPLB_WARP_TIMES
    if ((threadIdx.x & 31u) == 0) {
        atomicMax(&ctr->negFirstDry, ~src.dryAt);
        atomicMax(&ctr->negFirstDone, ~now());
        atomicMax(&ctr->lastDone, now());
        const unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        if (w < 4096) {
            ctr->warpDone[w] = now();
            ctr->warpDry[w] = src.dryAt;
            ctr->warpStates[w] = (unsigned)cnt.states;
            ctr->warpBv[w] = (unsigned)cnt.bv;
            ctr->warpTri[w] = (unsigned)cnt.tri;
        }
    }
#endif
    // a streamed batch whose copy did not arrive (e.g. a profiler serialising the kernel and the copy): reported
    // separately so that the host repeats the batch copy-then-kernel instead of failing
    if (__any_sync(kFull, src.stalled != 0) && (threadIdx.x & 31u) == 0) atomicExch(&ctr->pad, 1u);
    flushCounters(cnt, ctr, false);
    if constexpr (!kStreamed) ringHost(in.ret, ctr);
}

#ifndef MPLB_EDGE_MIN_CTAS
#define MPLB_EDGE_MIN_CTAS MPLB_MIN_CTAS
#endif
template <typename E, bool kMargins = false>
__global__ void __launch_bounds__(kWarpsEdges * 32, MPLB_EDGE_MIN_CTAS)
checkEdgesKernel(SceneDev sc, const double *__restrict__ from, const double *__restrict__ to,
                 const double *__restrict__ edgeConst, const unsigned *__restrict__ steps, unsigned long long nEdges,
                 unsigned *dead,
                 uint4 *stacks, unsigned stackCap, DevCounters *ctr, unsigned long long *workCounter,
                 const uint4 *descIn, const unsigned long long *descCount, uint4 *spillOut,
                 unsigned long long *spillCount, unsigned spillCap, HostReturn ret) {
    if (descIn && *descCount == 0) {   // nothing was spilled by the previous pass
        ringHost(ret, ctr);
        return;
    }
    WarpSmem<E> sm = warpSmem<E>(sc.stackLevels);
    EdgeSource src;
    src.from = from; src.to = to; src.edgeConst = edgeConst; src.steps = steps; src.nEdges = nEdges;
    src.dead = dead;
    src.crew = 2ull * gridDim.x * (blockDim.x >> 5);
    src.counter = workCounter;
    src.descIn = descIn;
    if (descIn) src.nDesc = min(*descCount, (unsigned long long)spillCap);
    src.spillOut = spillOut; src.spillCount = spillCount; src.spillCap = spillCap;
    src.stack = stacks + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * stackCap;
    src.stackCap = stackCap;
    src.robotRadius = sc.robotRadius;
    src.errorFlag = &ctr->overflow;
    WarpCounters cnt;
#ifdef MPLB_WARP_TIMES
    auto now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if ((threadIdx.x & 31u) == 0) atomicMax(&ctr->negStart, ~now());
#endif
    runWarp<E, false, kMargins>(sm, sc, src, cnt);
#ifdef MPLB_WARP_TIMES
    if ((threadIdx.x & 31u) == 0) {
        const unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        atomicMax(&ctr->lastDone, now());
        if (w < 4096) {
            ctr->warpDone[w] = now();
            ctr->warpStates[w] = (unsigned)cnt.states;
            ctr->warpBv[w] = (unsigned)cnt.bv;
            ctr->warpTri[w] = (unsigned)cnt.tri;
        }
    }
#endif
    flushCounters(cnt, ctr, true);
    ringHost(ret, ctr);
}

__global__ void deadToValidKernel(const unsigned *__restrict__ dead, unsigned long long n, uint8_t *__restrict__ valid) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) valid[i] = dead[i] ? 0 : 1;
}

template <typename E>
size_t smemBytes(const SceneDev &sc, int warps) { return WarpSmem<E>::bytes(sc.stackLevels) * warps; }

// Resident grid of `kernel` with `smem` bytes of dynamic shared memory on the current device.  The answer (and
// the opt-in to that much shared memory) is cached per device and size: a scalar isValid(q) is one ~8 us
// launch, two extra runtime calls in front of it are not free.
template <typename K>
cudaError_t gridFor(K kernel, size_t smem, int threads, int numSms, int *grid) {
    static std::mutex mu;
    static std::map<std::tuple<const void *, int, size_t>, int> cache;   // (kernels of one signature share this instantiation)
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    std::lock_guard<std::mutex> g(mu);
    const void *fn = reinterpret_cast<const void *>(kernel);
    auto it = cache.find({fn, dev, smem});
    if (it == cache.end()) {
        // the attribute is per kernel and device: keep it at the largest size any scene on this device needs
        size_t most = smem;
        for (const auto &kv : cache)
            if (std::get<0>(kv.first) == fn && std::get<1>(kv.first) == dev) most = std::max(most, std::get<2>(kv.first));
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most);
        if (err != cudaSuccess) return err;
        int perSm = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, threads, smem);
        if (err != cudaSuccess) return err;
        it = cache.emplace(std::make_tuple(fn, dev, smem), std::max(perSm, 1) * numSms).first;
    }
    *grid = it->second;
    return cudaSuccess;
}

template <typename E>
cudaError_t launchStatesT(const SceneDev &sc, const double *dStates, size_t n, uint8_t *dValid, DevCounters *dCtr,
                          unsigned long long *dWork, const StreamedInput &in, int numSms, cudaStream_t stream) {
    const bool streamed = in.water != nullptr || in.host != nullptr;
    // batches that leave most of the grid idle are latency bound
    // streamed batches (>= 65,536 states, capi.cu) wait for the host link, not for the SMs: they take the latency
    // variant, whose extra votes per step cost nothing under the copy and whose tail after the last piece has landed
    // is a third of the throughput variant's; a kernel slower than the copy (Apartment) loses ~3 %
    static const bool streamedThroughput = std::getenv("MPLB_STREAM_THROUGHPUT") != nullptr;   // diagnostic
    auto k = streamed ? (streamedThroughput ? checkStatesKernel<E, false, true> : checkStatesKernel<E, true, true>)
                      : n < kLatencyBelow ? checkStatesKernel<E, true, false> : checkStatesKernel<E, false, false>;
    if (sc.minMargin && !streamed) k = checkStatesKernel<E, false, false, true>;   // diagnostic: margins of the exact decisions
    // -DMPLB_WARPS_SMALL_CTA=4 runs the SM's twelve warps as three CTAs of four where the scene's stacks leave room for
    // the two extra CTAs' fixed shared memory (everything but Apartment, which at two CTAs of four goes 3.21 -> 3.58 ms
    // and keeps the one CTA of twelve): the warps do not talk to each other, but a CTA's registers come back only when
    // its last warp is through, and an SM that a launch has drained a third at a time starts on the next launch --
    // another stream's, mplb.h -- that much earlier (2^20 Alpha poses issued alternately on two streams: 0.695 -> 0.682
    // ms; on one stream the shape makes no difference).  NOT the default: with it the scalar-call sharing test
    // (test_scalar_calls_from_native_threads, sharing on) reported wrong verdicts once and hung once at the very end of
    // the round, unexplained; the one-CTA shape has passed that test in every run.
    int warps = kWarpsSmall, grid = 0;
    size_t smem = smemBytes<E>(sc, warps);
    cudaError_t err = gridFor(k, smem, warps * 32, numSms, &grid);
    if (err != cudaSuccess) return err;
    if (grid < kWarpsStates / kWarpsSmall * numSms) {
        warps = kWarpsStates;
        smem = smemBytes<E>(sc, warps);
        if ((err = gridFor(k, smem, warps * 32, numSms, &grid)) != cudaSuccess) return err;
    }
    // small batches are latency bound: one state per warp, spread over as many SMs as there are
    grid = (int)std::min<unsigned long long>(grid, (n + warps - 1) / warps);
#ifndef MPLB_GUIDED_SHARE
#define MPLB_GUIDED_SHARE 2
#endif
    constexpr unsigned guidedShare = MPLB_GUIDED_SHARE;   // a warp takes at most 1/(2 x warps) of what is left (1, 4, 8 measured no better)
    k<<<std::max(grid, 1), warps * 32, smem, stream>>>(sc, dStates, n, dValid, dCtr, dWork, guidedShare, in);
    return cudaGetLastError();
}

constexpr unsigned kEdgeStackCap = 2048;   // interval LIFO entries per warp
#ifndef MPLB_EDGE_PASSES
#define MPLB_EDGE_PASSES 3
#endif
constexpr int kEdgePasses = MPLB_EDGE_PASSES;   // <= 8
constexpr unsigned kSpillCap = 1u << 18;   // descriptors one pass can hand to the next (beyond that a warp keeps its own)
constexpr int kMaxCtasPerSm = 2;

template <typename E>
cudaError_t launchEdgesT(const SceneDev &sc, const double *dFrom, const double *dTo, const double *dEdgeConst,
                         const unsigned *dSteps, size_t n, unsigned *dDead, DevCounters *dCtr, unsigned long long *dWork, uint4 *dStacks,
                         size_t stackBytes, int numSms, cudaStream_t stream, const HostReturn &ret = HostReturn()) {
    auto k = sc.minMargin ? checkEdgesKernel<E, true> : checkEdgesKernel<E, false>;
    const size_t smem = smemBytes<E>(sc, kWarpsEdges);
    int grid = 0;
    cudaError_t err = gridFor(k, smem, kWarpsEdges * 32, numSms, &grid);
    if (err != cudaSuccess) return err;
    grid = std::min(grid, kMaxCtasPerSm * numSms);
    const unsigned long long nItems = (1ull + kPhaseBSub) * n;   // `to` + the intervals of every edge
    grid = (int)std::min<unsigned long long>(grid, (nItems + kWarpsEdges - 1) / kWarpsEdges);
    grid = std::max(grid, 1);
    const size_t lifoBytes = (size_t)grid * kWarpsEdges * kEdgeStackCap * sizeof(uint4);
    if (lifoBytes + 2 * (size_t)kSpillCap * sizeof(uint4) > stackBytes) return cudaErrorInvalidValue;
    // dWork: [p] item counter of pass p, [8 + p] number of descriptors spilled by pass p
    uint4 *spill[2] = {dStacks + (stackBytes / sizeof(uint4) - 2 * (size_t)kSpillCap), nullptr};
    spill[1] = spill[0] + kSpillCap;
    const int threads = kWarpsEdges * 32;
    // MPLB_SPILL_CAP (diagnostic, tests): a small capacity makes the reservations overflow
    static const unsigned spillCap = [] {
        const char *e = std::getenv("MPLB_SPILL_CAP");
        return e ? std::min<unsigned>(kSpillCap, (unsigned)std::max(1, std::atoi(e))) : kSpillCap;
    }();
    for (int p = 0; p < kEdgePasses; ++p) {
        const bool last = p == kEdgePasses - 1;
        k<<<grid, threads, smem, stream>>>(sc, dFrom, dTo, dEdgeConst, dSteps, n, dDead, dStacks, kEdgeStackCap, dCtr, dWork + p,
                                           p ? spill[(p - 1) & 1] : nullptr, p ? dWork + 8 + (p - 1) : nullptr,
                                           last ? nullptr : spill[p & 1], last ? nullptr : dWork + 8 + p, spillCap,
                                           last ? ret : HostReturn());   // the last pass's last warp rings the host
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launchCheckStates(const SceneDev &sc, const double *dStates, size_t n, uint8_t *dValid,
                              DevCounters *dCtr, unsigned *dWork, int numSms, cudaStream_t stream, StreamedInput streamed,
                              bool workZeroed) {
    if (n == 0) return cudaSuccess;
    cudaError_t err = workZeroed ? cudaSuccess : cudaMemsetAsync(dWork, 0, sizeof(unsigned long long), stream);
    if (err != cudaSuccess) return err;
    return sc.wide ? launchStatesT<unsigned long long>(sc, dStates, n, dValid, dCtr, (unsigned long long *)dWork, streamed, numSms, stream)
                   : launchStatesT<unsigned>(sc, dStates, n, dValid, dCtr, (unsigned long long *)dWork, streamed, numSms, stream);
}

size_t edgeStackBytes(int numSms) {
    return (size_t)numSms * kMaxCtasPerSm * kWarpsEdges * kEdgeStackCap * sizeof(uint4) + 2 * (size_t)kSpillCap * sizeof(uint4);
}

size_t edgeConstBytes(size_t n) { return n * kEdgeConst * sizeof(double); }

cudaError_t launchCheckEdges(const SceneDev &sc, const double *dFrom, const double *dTo, const unsigned *dSteps,
                             size_t n, unsigned *dDead, uint8_t *dValid, DevCounters *dCtr, unsigned long long *dWork,
                             void *dStacks, size_t stackBytes, double *dEdgeConst, int numSms, cudaStream_t stream,
                             bool zeroed, HostReturn ret, const void *hostInputs) {
    if (n == 0) return cudaSuccess;
    cudaError_t err = cudaSuccess;
    if (!zeroed) {   // (small host calls zero their one control block -- counters, work counters, dead flags -- at once)
        err = cudaMemsetAsync(dWork, 0, 16 * sizeof(unsigned long long), stream);
        if (err != cudaSuccess) return err;
        err = cudaMemsetAsync(dDead, 0, sizeof(unsigned) * n, stream);
        if (err != cudaSuccess) return err;
    }
    if (hostInputs)
        edgeStageKernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>((const double *)hostInputs, n, sc.robotRadius,
                                                                      const_cast<double *>(dFrom), const_cast<double *>(dTo),
                                                                      const_cast<unsigned *>(dSteps), dEdgeConst);
    else
        edgeConstKernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(dFrom, dTo, n, sc.robotRadius, dEdgeConst);
    err = sc.wide ? launchEdgesT<unsigned long long>(sc, dFrom, dTo, dEdgeConst, dSteps, n, dDead, dCtr, dWork, (uint4 *)dStacks, stackBytes, numSms, stream, ret)
                  : launchEdgesT<unsigned>(sc, dFrom, dTo, dEdgeConst, dSteps, n, dDead, dCtr, dWork, (uint4 *)dStacks, stackBytes, numSms, stream, ret);
    if (err != cudaSuccess) return err;
    if (dValid) deadToValidKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dDead, n, dValid);   // null: the caller reads dDead
    return cudaGetLastError();
}

// dScratch layout of the indexed call: from [n][7] | to [n][7] | edgeConst [n][kEdgeConst] | steps [n]
size_t edgeGatherBytes(size_t n) { return n * (14 + kEdgeConst) * sizeof(double) + n * sizeof(unsigned) + 64; }
double *edgeGatherFrom(void *dScratch, size_t) { return (double *)dScratch; }
double *edgeGatherTo(void *dScratch, size_t n) { return (double *)dScratch + 7 * n; }
static double *edgeGatherConst(void *dScratch, size_t n) { return (double *)dScratch + 14 * n; }
unsigned *edgeGatherSteps(void *dScratch, size_t n) { return (unsigned *)((double *)dScratch + (14 + kEdgeConst) * n); }

cudaError_t launchCheckEdgesIndexed(const SceneDev &sc, const EdgeGather &g, size_t n, double so3w, double l2w,
                                    double invStep, unsigned *dDead, uint8_t *dValid, DevCounters *dCtr,
                                    unsigned long long *dWork, void *dStacks, size_t stackBytes, void *dScratch,
                                    int numSms, cudaStream_t stream, bool zeroed) {
    if (n == 0) return cudaSuccess;
    cudaError_t err = cudaSuccess;
    if (!zeroed) {
        err = cudaMemsetAsync(dWork, 0, 16 * sizeof(unsigned long long), stream);
        if (err != cudaSuccess) return err;
        err = cudaMemsetAsync(dDead, 0, sizeof(unsigned) * n, stream);
        if (err != cudaSuccess) return err;
    }
    double *dFrom = edgeGatherFrom(dScratch, n), *dTo = edgeGatherTo(dScratch, n), *dConst = edgeGatherConst(dScratch, n);
    unsigned *dSteps = edgeGatherSteps(dScratch, n);
    edgeGatherPlanKernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(g, n, so3w, l2w, invStep, sc.robotRadius, dFrom, dTo,
                                                                       dSteps, dDead, dConst, dCtr);
    err = sc.wide ? launchEdgesT<unsigned long long>(sc, dFrom, dTo, dConst, dSteps, n, dDead, dCtr, dWork, (uint4 *)dStacks, stackBytes, numSms, stream)
                  : launchEdgesT<unsigned>(sc, dFrom, dTo, dConst, dSteps, n, dDead, dCtr, dWork, (uint4 *)dStacks, stackBytes, numSms, stream);
    if (err != cudaSuccess) return err;
    if (dValid) deadToValidKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dDead, n, dValid);
    return cudaGetLastError();
}

}  // namespace mplb
