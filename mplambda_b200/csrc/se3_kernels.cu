// SE(3) rigid-body validity kernels for sm_100a.
//
// Device replacement for the work behind
//   SE3RigidBodyScenario::isValid(q)        reference include/mpl/demo/se3_rigid_body_scenario.hpp:119-126
//   SE3RigidBodyScenario::isValid(from,to)  reference include/mpl/demo/se3_rigid_body_scenario.hpp:134-178
//   mpl::interpolate (slerp / lerp)         reference include/mpl/interpolate.hpp:9-36
// i.e. fcl::collide(robotMesh, tf(q), envMesh, Identity) per state ([EXT] FCL: OBB hierarchy
// traversal + 17-axis triangle-triangle separating-axis test, first contact wins).
//
// Execution model (one warp = one cooperative work pool):
//   * a warp owns a chunk of 64 states (64 consecutive poses, or 64 interpolated states of one
//     edge in bit-reversed -- coarse-to-fine -- order) and expands them to FP32 (R,T) slots in
//     shared memory;
//   * (slot, robotNode, envNode) pairs live in a per-warp LIFO in shared memory; every
//     iteration the 32 lanes pop 32 pairs -- of whatever states -- run one OBB-OBB SAT each
//     and push the surviving child pairs with a ballot/popc prefix, so lanes stay busy no
//     matter how uneven the per-state traversal depth is;
//   * leaf-leaf survivors go to a second queue and are tested 32 at a time with an
//     error-bounded FP32 17-axis SAT; the rare pairs whose FP32 verdict is not certain go to a
//     third queue and are re-evaluated in FP64 with exactly the oracle's operation order
//     (__dmul_rn/__dadd_rn: no FMA contraction), which is what makes verdicts bit-identical to
//     the FP64 CPU path instead of "equal up to an epsilon band";
//   * per-state hit bits are warp-uniform registers (ballot + reduce_or); for edges the first
//     hit abandons the chunk and marks the edge dead so later chunks of it are skipped.
// Node and triangle fetches are 128-bit read-only loads (4 per 64-byte node, 3 per triangle);
// the hierarchy is small enough to live in L1/L2, HBM only sees states
// in and verdict bytes out.
#include "se3_device.cuh"

#include <cfloat>

namespace mplb {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kChunk = 64;          // states per work item
constexpr int kStackCap = 1024;     // pair LIFO entries per warp
// Above this fill level a warp pops fewer pairs per iteration (down to one: plain depth-first
// search, which grows by at most one entry per level), so the LIFO cannot overflow as long as
// depthA + depthB <= kStackCap - kStackSoftCap (checked at scene creation).
constexpr int kStackSoftCap = kStackCap - 72;
constexpr int kQueueCap = 96;       // leaf / exact queue entries per warp
constexpr unsigned kFull = 0xffffffffu;

template <typename E>
struct WarpSmem {
    E stack[kStackCap];
    E triq[kQueueCap];
    E exq[kQueueCap];
    float4 xf[kChunk][4];  // per slot: rows of (R | T), then (slack, eta, -, -); see makeSlot
};

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// ---------------------------------------------------------------------------------------------
// FP64 "specification" arithmetic: identical operation order to oracle/mplb_oracle.c.

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3sub(D3 a, D3 b) { return {dsub(a.x, b.x), dsub(a.y, b.y), dsub(a.z, b.z)}; }
__device__ __forceinline__ double d3dot(D3 a, D3 b) {
    return dadd(dadd(dmul(a.x, b.x), dmul(a.y, b.y)), dmul(a.z, b.z));
}
__device__ __forceinline__ D3 d3cross(D3 a, D3 b) {
    return {dsub(dmul(a.y, b.z), dmul(a.z, b.y)), dsub(dmul(a.z, b.x), dmul(a.x, b.z)),
            dsub(dmul(a.x, b.y), dmul(a.y, b.x))};
}

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

// interpolate.hpp:18-36 with Eigen's slerp; same statements as orc_se3_interpolate.  sin/acos are
// CUDA's FP64 libdevice (<= 2 ulp from glibc's, see DESIGN.md "numerics").
__device__ void interpolateD(const double *from, const double *to, double t, double out[7]) {
    double d = dadd(dadd(dadd(dmul(from[0], to[0]), dmul(from[1], to[1])), dmul(from[2], to[2])),
                    dmul(from[3], to[3]));
    double absD = fabs(d), s0, s1;
    if (absD >= 1.0 - DBL_EPSILON) {
        s0 = dsub(1.0, t);
        s1 = t;
    } else {
        double theta = acos(absD), sinTheta = sin(theta);
        s0 = __ddiv_rn(sin(dmul(dsub(1.0, t), theta)), sinTheta);
        s1 = __ddiv_rn(sin(dmul(t, theta)), sinTheta);
    }
    if (d < 0) s1 = -s1;
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = dadd(dmul(s0, from[i]), dmul(s1, to[i]));
#pragma unroll
    for (int i = 4; i < 7; ++i) out[i] = dadd(from[i], dmul(dsub(to[i], from[i]), t));
}

// ---------------------------------------------------------------------------------------------
// FP32 fast path

// Round the FP64 relative transform to the FP32 slot and derive the error parameters.
__device__ void makeSlot(const double s[7], const SceneDev &sc, float4 *xf) {
    XformD X;
    stateToRelD(s, X);
    double qn = s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + s[3] * s[3];
    double scale = fmax(1.0, qn);  // |R| for non-unit quaternions
    double tn = sqrt(s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
    double mag = scale * ((double)sc.envRadius + tn) + (double)sc.robotRadius;
    float eta = (float)(8.0 * FLT_EPSILON * mag) * 1.0000002f;
    float slack = (float)(64.0 * FLT_EPSILON * (mag + 4.0 * scale * (double)sc.envRadius +
                                                4.0 * (double)sc.robotRadius));
    xf[0] = make_float4((float)X.R[0][0], (float)X.R[0][1], (float)X.R[0][2], (float)X.T[0]);
    xf[1] = make_float4((float)X.R[1][0], (float)X.R[1][1], (float)X.R[1][2], (float)X.T[1]);
    xf[2] = make_float4((float)X.R[2][0], (float)X.R[2][1], (float)X.R[2][2], (float)X.T[2]);
    xf[3] = make_float4(slack, eta, 0.f, 0.f);
}

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float dot(F3 a, F3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ F3 cross(F3 a, F3 b) {
    return {fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x)};
}

// OBB-OBB separating-axis test (Gottschalk's 15 axes).  Conservative: `slack` exceeds the FP32
// evaluation error of either side, so a pair is only culled when the exact boxes are disjoint.
// A = robot node (robot frame), B = env node (env frame), (R,T) maps env -> robot frame.
__device__ __forceinline__ bool obbOverlap(const float4 a0, const float4 a1, const float4 a2, const float4 ae,
                                           const float4 b0, const float4 b1, const float4 b2, const float4 be,
                                           const float4 x0, const float4 x1, const float4 x2, float slack) {
    // env centre in the robot frame, relative to the robot box centre, in the robot box axes
    float cx = fmaf(x0.x, b0.w, fmaf(x0.y, b1.w, fmaf(x0.z, b2.w, x0.w))) - a0.w;
    float cy = fmaf(x1.x, b0.w, fmaf(x1.y, b1.w, fmaf(x1.z, b2.w, x1.w))) - a1.w;
    float cz = fmaf(x2.x, b0.w, fmaf(x2.y, b1.w, fmaf(x2.z, b2.w, x2.w))) - a2.w;
    float T[3] = {fmaf(a0.x, cx, fmaf(a0.y, cy, a0.z * cz)), fmaf(a1.x, cx, fmaf(a1.y, cy, a1.z * cz)),
                  fmaf(a2.x, cx, fmaf(a2.y, cy, a2.z * cz))};
    // RB_j = R * b_j
    F3 rb[3];
    {
        const float4 b[3] = {b0, b1, b2};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            rb[j].x = fmaf(x0.x, b[j].x, fmaf(x0.y, b[j].y, x0.z * b[j].z));
            rb[j].y = fmaf(x1.x, b[j].x, fmaf(x1.y, b[j].y, x1.z * b[j].z));
            rb[j].z = fmaf(x2.x, b[j].x, fmaf(x2.y, b[j].y, x2.z * b[j].z));
        }
    }
    float B[3][3], Bf[3][3];
    {
        const float4 a[3] = {a0, a1, a2};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                B[i][j] = fmaf(a[i].x, rb[j].x, fmaf(a[i].y, rb[j].y, a[i].z * rb[j].z));
                Bf[i][j] = fabsf(B[i][j]);
            }
    }
    const float ea[3] = {ae.x, ae.y, ae.z}, eb[3] = {be.x, be.y, be.z};
    bool sep = false;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        sep |= fabsf(T[i]) > ea[i] + fmaf(eb[0], Bf[i][0], fmaf(eb[1], Bf[i][1], fmaf(eb[2], Bf[i][2], slack)));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float s = fmaf(T[0], B[0][j], fmaf(T[1], B[1][j], T[2] * B[2][j]));
        sep |= fabsf(s) > eb[j] + fmaf(ea[0], Bf[0][j], fmaf(ea[1], Bf[1][j], fmaf(ea[2], Bf[2][j], slack)));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            float s = fmaf(T[i2], B[i1][j], -T[i1] * B[i2][j]);
            float r = fmaf(ea[i1], Bf[i2][j], fmaf(ea[i2], Bf[i1][j], fmaf(eb[j1], Bf[i][j2], fmaf(eb[j2], Bf[i][j1], slack))));
            sep |= fabsf(s) > r;
        }
    }
    return !sep;
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
    const F3 e[3] = {p2, p3 - p2, f3(-p3.x, -p3.y, -p3.z)};
    const F3 f[3] = {f1, f2 - f1, f3(-f2.x, -f2.y, -f2.z)};
    const float Me = fmaxf(absMax(e[0]), fmaxf(absMax(e[1]), absMax(e[2])));
    const float Mf = fmaxf(absMax(f[0]), fmaxf(absMax(f[1]), absMax(f[2])));
    const float U = fmaxf(fmaxf(fmaxf(absMax(p2), absMax(p3)), fmaxf(absMax(q1), absMax(q2))), absMax(q3));
    const float S = fmaf(79.f * FLT_EPSILON, U, eta);
    const float eu = FLT_EPSILON * U;
    const float Kn = 120.f * eu * Me * Me, Km = 792.f * eu * Mf * Mf, Kef = 456.f * eu * Me * Mf;
    const float Kg = 384.f * eu * Me * Me * Me, Kh = 2400.f * eu * Mf * Mf * Mf;
    const F3 n1 = cross(e[0], e[1]), m1 = cross(f[0], f[1]);
    AxisState st{false, false};
    axisTest(n1, p2, p3, q1, q2, q3, S, Kn, st);
    axisTest(m1, p2, p3, q1, q2, q3, S, Km, st);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) axisTest(cross(e[i], f[j]), p2, p3, q1, q2, q3, S, Kef, st);
#pragma unroll
    for (int i = 0; i < 3; ++i) axisTest(cross(e[i], n1), p2, p3, q1, q2, q3, S, Kg, st);
#pragma unroll
    for (int i = 0; i < 3; ++i) axisTest(cross(f[i], m1), p2, p3, q1, q2, q3, S, Kh, st);
    if (st.sep) return kTriNo;
    return st.unsure ? kTriUnsure : kTriYes;
}

// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ unsigned bitrev(unsigned j, int bits) { return bits ? (__brev(j) >> (32 - bits)) : 0u; }

template <typename E>
struct Codec {
    int shA, shB;  // entry = slot | a << 6 | b << shB
    __device__ __forceinline__ E pack(unsigned slot, unsigned a, unsigned b) const {
        return (E)slot | ((E)a << 6) | ((E)b << shB);
    }
    __device__ __forceinline__ unsigned slot(E e) const { return (unsigned)(e & 63); }
    __device__ __forceinline__ unsigned a(E e) const { return (unsigned)((e >> 6) & (((E)1 << (shB - 6)) - 1)); }
    __device__ __forceinline__ unsigned b(E e) const { return (unsigned)(e >> shB); }
};

// One warp drains its pool: returns the 64-bit hit mask of the chunk (bit s = slot s collides).
// EDGE: stop at the first hit (the whole chunk belongs to one edge).
template <typename E, bool EDGE, typename StateFn>
__device__ unsigned long long drainPool(WarpSmem<E> &sm, const SceneDev &sc, const Codec<E> cd, int top,
                                        StateFn &&stateOf, unsigned long long &nBv, unsigned long long &nTri,
                                        unsigned long long &nExact, unsigned &overflow) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned ltMask = (1u << lane) - 1u;
    unsigned hitLo = 0, hitHi = 0;
    int nt = 0, nx = 0;
    for (;;) {
        if (EDGE && (hitLo | hitHi)) break;
        if (nt >= sc.triBatchMin || (top < 32 && nt > 0)) {
            // ---- triangle-triangle batch (FP32 filter)
            const int n = min(32, nt);
            const bool act0 = (int)lane < n;
            E en = act0 ? sm.triq[nt - 1 - lane] : 0;
            nt -= n;
            __syncwarp();
            const unsigned slot = cd.slot(en);
            const bool dead = slot < 32 ? (hitLo >> slot) & 1u : (hitHi >> (slot - 32)) & 1u;
            const bool act = act0 && !dead;
            int verdict = kTriNo;
            if (act) {
                const float4 *tp = sc.robotTris + 3 * (size_t)cd.a(en);
                const float4 *tq = sc.envTris + 3 * (size_t)cd.b(en);
                const float4 p1 = ldg4(tp), p2 = ldg4(tp + 1), p3 = ldg4(tp + 2);
                const float4 w1 = ldg4(tq), w2 = ldg4(tq + 1), w3 = ldg4(tq + 2);
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
            nTri += __popc(__ballot_sync(kFull, act));
            hitLo |= __reduce_or_sync(kFull, (verdict == kTriYes && slot < 32) ? 1u << slot : 0u);
            hitHi |= __reduce_or_sync(kFull, (verdict == kTriYes && slot >= 32) ? 1u << (slot - 32) : 0u);
            const unsigned mu = __ballot_sync(kFull, verdict == kTriUnsure);
            if (mu) {
                if (verdict == kTriUnsure) sm.exq[nx + __popc(mu & ltMask)] = en;
                nx += __popc(mu);
                __syncwarp();
            }
            if (nx < 32) continue;
        }
        if (nx >= 32 || (top < 32 && nt == 0 && nx > 0)) {
            // ---- exact FP64 batch
            const int n = min(32, nx);
            const bool act0 = (int)lane < n;
            E en = act0 ? sm.exq[nx - 1 - lane] : 0;
            nx -= n;
            __syncwarp();
            const unsigned slot = cd.slot(en);
            const bool dead = slot < 32 ? (hitLo >> slot) & 1u : (hitHi >> (slot - 32)) & 1u;
            const bool act = act0 && !dead;
            bool hit = false;
            if (act) {
                double s[7];
                stateOf(slot, s);
                XformD X;
                stateToRelD(s, X);
                hit = triTriExact(sc.robotTris64 + 9 * (size_t)cd.a(en), sc.envTris64 + 9 * (size_t)cd.b(en), X);
            }
            nExact += __popc(__ballot_sync(kFull, act));
            hitLo |= __reduce_or_sync(kFull, (hit && slot < 32) ? 1u << slot : 0u);
            hitHi |= __reduce_or_sync(kFull, (hit && slot >= 32) ? 1u << (slot - 32) : 0u);
            continue;
        }
        if (top == 0) break;
        // ---- OBB-OBB batch
        const int n = min(min(32, top), max(1, kStackSoftCap - top));
        const bool act0 = (int)lane < n;
        E en = act0 ? sm.stack[top - 1 - lane] : 0;
        top -= n;
        __syncwarp();
        const unsigned slot = cd.slot(en);
        const bool dead = slot < 32 ? (hitLo >> slot) & 1u : (hitHi >> (slot - 32)) & 1u;
        const bool act = act0 && !dead;
        bool ov = false, leafA = false, leafB = false, descendA = false;
        int linkA = 0, linkB = 0;
        const unsigned na = cd.a(en), nb = cd.b(en);
        if (act) {
            const float4 *pa = sc.robotNodes + 4 * (size_t)na;
            const float4 *pb = sc.envNodes + 4 * (size_t)nb;
            const float4 a0 = ldg4(pa), a1 = ldg4(pa + 1), a2 = ldg4(pa + 2), ae = ldg4(pa + 3);
            const float4 b0 = ldg4(pb), b1 = ldg4(pb + 1), b2 = ldg4(pb + 2), be = ldg4(pb + 3);
            const float4 x0 = sm.xf[slot][0], x1 = sm.xf[slot][1], x2 = sm.xf[slot][2], x3 = sm.xf[slot][3];
            ov = obbOverlap(a0, a1, a2, ae, b0, b1, b2, be, x0, x1, x2, x3.x);
            linkA = __float_as_int(ae.w);
            linkB = __float_as_int(be.w);
            leafA = linkA < 0;
            leafB = linkB < 0;
            const float szA = fmaf(ae.x, ae.x, fmaf(ae.y, ae.y, ae.z * ae.z));
            const float szB = fmaf(be.x, be.x, fmaf(be.y, be.y, be.z * be.z));
            descendA = leafB || (!leafA && szA > szB);
        }
        nBv += __popc(__ballot_sync(kFull, act));
        const bool toTri = ov && leafA && leafB;
        const bool toBv = ov && !toTri;
        const unsigned mt = __ballot_sync(kFull, toTri);
        const unsigned mb = __ballot_sync(kFull, toBv);
        if (toTri) sm.triq[nt + __popc(mt & ltMask)] = cd.pack(slot, (unsigned)(-(linkA + 1)), (unsigned)(-(linkB + 1)));
        nt += __popc(mt);
        if (top + 2 * __popc(mb) > kStackCap) {
            overflow = 1;  // cannot happen with the builder's depth bound; reported, never ignored
            break;
        }
        if (toBv) {
            const int pos = top + 2 * __popc(mb & ltMask);
            if (descendA) {
                sm.stack[pos] = cd.pack(slot, (unsigned)linkA + 1, nb);
                sm.stack[pos + 1] = cd.pack(slot, (unsigned)linkA, nb);
            } else {
                sm.stack[pos] = cd.pack(slot, na, (unsigned)linkB + 1);
                sm.stack[pos + 1] = cd.pack(slot, na, (unsigned)linkB);
            }
        }
        top += 2 * __popc(mb);
        __syncwarp();
    }
    return ((unsigned long long)hitHi << 32) | hitLo;
}

template <typename E>
__device__ __forceinline__ WarpSmem<E> &warpSmem() {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    return reinterpret_cast<WarpSmem<E> *>(smemRaw)[threadIdx.x >> 5];
}

// isValid(q) for a batch: one work item = 64 consecutive states.
template <typename E>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
checkStatesKernel(SceneDev sc, const double *__restrict__ states, unsigned long long n, uint8_t *__restrict__ valid,
                  DevCounters *ctr, unsigned *workCounter) {
    WarpSmem<E> &sm = warpSmem<E>();
    const unsigned lane = threadIdx.x & 31u;
    const Codec<E> cd{6, 6 + sc.bitsA};
    const unsigned long long nChunks = (n + kChunk - 1) / kChunk;
    unsigned long long nBv = 0, nTri = 0, nExact = 0;
    unsigned overflow = 0;
    const bool haveGeom = sc.nRobotNodes > 0 && sc.nEnvNodes > 0;
    for (;;) {
        unsigned long long ck = 0;
        if (lane == 0) ck = atomicAdd(workCounter, 1u);
        ck = __shfl_sync(kFull, ck, 0);
        if (ck >= nChunks) break;
        const unsigned long long base = ck * kChunk;
        int top = 0;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const unsigned slot = r * 32 + lane;
            const bool ok = base + slot < n;
            if (ok && haveGeom) {
                double s[7];
                const double *src = states + 7 * (base + slot);
#pragma unroll
                for (int k = 0; k < 7; ++k) s[k] = __ldg(src + k);
                makeSlot(s, sc, sm.xf[slot]);
            }
            const unsigned m = __ballot_sync(kFull, ok && haveGeom);
            if (ok && haveGeom) sm.stack[top + __popc(m & ((1u << lane) - 1u))] = cd.pack(slot, 0, 0);
            top += __popc(m);
        }
        __syncwarp();
        auto stateOf = [&](unsigned slot, double *s) {
            const double *src = states + 7 * (base + slot);
#pragma unroll
            for (int k = 0; k < 7; ++k) s[k] = __ldg(src + k);
        };
        const unsigned long long hits = drainPool<E, false>(sm, sc, cd, top, stateOf, nBv, nTri, nExact, overflow);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const unsigned slot = r * 32 + lane;
            if (base + slot < n) valid[base + slot] = (uint8_t)(((hits >> slot) & 1ull) ? 0 : 1);
        }
        __syncwarp();
    }
    if (lane == 0) {
        atomicAdd(&ctr->bvTests, nBv);
        atomicAdd(&ctr->triTests, nTri);
        atomicAdd(&ctr->exactTests, nExact);
        if (overflow) atomicExch(&ctr->overflow, 1u);
    }
}

// isValid(from,to) for a batch of edges.  Work items are (chunk c, edge e), c-major: chunk 0 of
// every edge (the coarsest 64 samples incl. `to`) runs before any chunk 1, so an edge that dies
// early never expands its fine samples.  States of chunk c: j = 64c + slot, i = bitrev_L(j)
// with 2^L > steps; j = 0 is `to` itself (se3...:136), i in [1, steps-1] are the interpolated
// states at t = i * (1/steps) (se3...:146,159).
template <typename E>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
checkEdgesKernel(SceneDev sc, const double *__restrict__ from, const double *__restrict__ to,
                 const unsigned *__restrict__ steps, unsigned long long nEdges, unsigned maxChunks, unsigned grab,
                 unsigned *dead, DevCounters *ctr, unsigned long long *workCounter) {
    WarpSmem<E> &sm = warpSmem<E>();
    const unsigned lane = threadIdx.x & 31u;
    const Codec<E> cd{6, 6 + sc.bitsA};
    const unsigned long long nItems = nEdges * maxChunks;
    unsigned long long nBv = 0, nTri = 0, nExact = 0, nStates = 0;
    unsigned overflow = 0;
    const bool haveGeom = sc.nRobotNodes > 0 && sc.nEnvNodes > 0;
    for (;;) {
        unsigned long long first = 0;
        if (lane == 0) first = atomicAdd(workCounter, (unsigned long long)grab);
        first = __shfl_sync(kFull, first, 0);
        if (first >= nItems) break;
        // each lane inspects one item of the grab; live ones are processed one after another
        const unsigned long long item = first + lane;
        bool live = false;
        unsigned myBits = 0;
        if (lane < grab && item < nItems && haveGeom) {
            const unsigned long long e = item % nEdges;
            const unsigned c = (unsigned)(item / nEdges);
            const unsigned st = __ldg(steps + e);
            myBits = st ? 32 - __clz(st) : 0;  // 2^bits > steps
            const unsigned nch = max(1u, (1u << myBits) >> 6);
            live = c < nch && !((volatile unsigned *)dead)[e];
        }
        unsigned liveMask = __ballot_sync(kFull, live);
        while (liveMask) {
            const int src = __ffs(liveMask) - 1;
            liveMask &= liveMask - 1;
            const unsigned long long it = first + src;
            const unsigned long long e = it % nEdges;
            const unsigned c = (unsigned)(it / nEdges);
            const unsigned bits = __shfl_sync(kFull, myBits, src);
            const unsigned st = __ldg(steps + e);
            const double *fr = from + 7 * e, *tt = to + 7 * e;
            double a[7], b[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) { a[k] = __ldg(fr + k); b[k] = __ldg(tt + k); }
            const double delta = __ddiv_rn(1.0, (double)st);
            auto stateOf = [&](unsigned slot, double *s) {
                const unsigned j = c * kChunk + slot;
                if (j == 0) {
#pragma unroll
                    for (int k = 0; k < 7; ++k) s[k] = b[k];
                } else {
                    interpolateD(a, b, dmul((double)bitrev(j, bits), delta), s);
                }
            };
            int top = 0;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const unsigned slot = r * 32 + lane;
                const unsigned j = c * kChunk + slot;
                const bool ok = j == 0 || (j < (1u << bits) && bitrev(j, bits) < st);
                if (ok) {
                    double s[7];
                    stateOf(slot, s);
                    makeSlot(s, sc, sm.xf[slot]);
                }
                const unsigned m = __ballot_sync(kFull, ok);
                if (ok) sm.stack[top + __popc(m & ((1u << lane) - 1u))] = cd.pack(slot, 0, 0);
                top += __popc(m);
            }
            __syncwarp();
            nStates += top;
            const unsigned long long hits = drainPool<E, true>(sm, sc, cd, top, stateOf, nBv, nTri, nExact, overflow);
            if (hits && lane == 0) dead[e] = 1u;
            __syncwarp();
        }
    }
    if (lane == 0) {
        atomicAdd(&ctr->bvTests, nBv);
        atomicAdd(&ctr->triTests, nTri);
        atomicAdd(&ctr->exactTests, nExact);
        atomicAdd(&ctr->states, nStates);
        if (overflow) atomicExch(&ctr->overflow, 1u);
    }
}

__global__ void deadToValidKernel(const unsigned *__restrict__ dead, unsigned long long n, uint8_t *__restrict__ valid) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) valid[i] = dead[i] ? 0 : 1;
}

template <typename E>
size_t smemBytes() { return sizeof(WarpSmem<E>) * kWarpsPerCta; }

template <typename K>
int gridFor(K kernel, size_t smem, int numSms) {
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kWarpsPerCta * 32, smem);
    if (perSm < 1) perSm = 1;
    return perSm * numSms;
}

}  // namespace

cudaError_t launchCheckStates(const SceneDev &sc, const double *dStates, size_t n, uint8_t *dValid,
                              DevCounters *dCtr, unsigned *dWork, int numSms, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    cudaError_t err = cudaMemsetAsync(dWork, 0, sizeof(unsigned long long), stream);
    if (err != cudaSuccess) return err;
    const unsigned long long nChunks = (n + kChunk - 1) / kChunk;
    if (sc.wide) {
        auto k = checkStatesKernel<unsigned long long>;
        const size_t smem = smemBytes<unsigned long long>();
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int grid = gridFor(k, smem, numSms);
        grid = (int)min((unsigned long long)grid, (nChunks + kWarpsPerCta - 1) / kWarpsPerCta);
        k<<<grid, kWarpsPerCta * 32, smem, stream>>>(sc, dStates, n, dValid, dCtr, dWork);
    } else {
        auto k = checkStatesKernel<unsigned>;
        const size_t smem = smemBytes<unsigned>();
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int grid = gridFor(k, smem, numSms);
        grid = (int)min((unsigned long long)grid, (nChunks + kWarpsPerCta - 1) / kWarpsPerCta);
        k<<<grid, kWarpsPerCta * 32, smem, stream>>>(sc, dStates, n, dValid, dCtr, dWork);
    }
    return cudaGetLastError();
}

cudaError_t launchCheckEdges(const SceneDev &sc, const double *dFrom, const double *dTo, const unsigned *dSteps,
                             size_t n, unsigned maxSteps, unsigned *dDead, uint8_t *dValid, DevCounters *dCtr,
                             unsigned long long *dWork, int numSms, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    cudaError_t err = cudaMemsetAsync(dWork, 0, sizeof(unsigned long long), stream);
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(dDead, 0, sizeof(unsigned) * n, stream);
    if (err != cudaSuccess) return err;
    const unsigned bits = maxSteps ? 32 - __builtin_clz(maxSteps) : 0;
    const unsigned maxChunks = std::max(1u, (1u << bits) >> 6);
    const unsigned long long nItems = (unsigned long long)n * maxChunks;
    int grid;
    unsigned grab;
    if (sc.wide) {
        auto k = checkEdgesKernel<unsigned long long>;
        const size_t smem = smemBytes<unsigned long long>();
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        grid = gridFor(k, smem, numSms);
        grab = (unsigned)std::min<unsigned long long>(32, std::max<unsigned long long>(1, nItems / (4ull * grid * kWarpsPerCta)));
        grid = (int)std::min<unsigned long long>(grid, (nItems + kWarpsPerCta - 1) / kWarpsPerCta);
        k<<<grid, kWarpsPerCta * 32, smem, stream>>>(sc, dFrom, dTo, dSteps, n, maxChunks, grab, dDead, dCtr, dWork);
    } else {
        auto k = checkEdgesKernel<unsigned>;
        const size_t smem = smemBytes<unsigned>();
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        grid = gridFor(k, smem, numSms);
        grab = (unsigned)std::min<unsigned long long>(32, std::max<unsigned long long>(1, nItems / (4ull * grid * kWarpsPerCta)));
        grid = (int)std::min<unsigned long long>(grid, (nItems + kWarpsPerCta - 1) / kWarpsPerCta);
        k<<<grid, kWarpsPerCta * 32, smem, stream>>>(sc, dFrom, dTo, dSteps, n, maxChunks, grab, dDead, dCtr, dWork);
    }
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    deadToValidKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dDead, n, dValid);
    return cudaGetLastError();
}

}  // namespace mplb
