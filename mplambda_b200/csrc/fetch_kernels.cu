// Fetch 8-DOF arm validity kernels for sm_100a.
//
// Device replacement for the work behind FetchScenario::isValid (reference
// include/mpl/demo/fetch_scenario.hpp:105-173):
//   FetchRobot::fk() + tf*Link()      include/mpl/demo/fetch_robot.hpp:217-315   (12 geometry frames)
//   FetchRobot::selfCollision()       include/mpl/demo/fetch_robot.hpp:502-584   (24 ordered shape pairs + floor)
//   FetchRobot::inCollisionWith()     include/mpl/demo/fetch_robot.hpp:586-647   (12 shapes vs the environment mesh)
// whose arithmetic is FCL + libccd ([EXT]): shape pairs go through libccd's MPR boolean test
// (ccdMPRIntersect) with FCL's support functions, Box x Box through FCL's boxBox2, mesh-vs-shape
// through an OBB hierarchy traversal with an MPR shape-triangle leaf test.
//
// Everything that decides a verdict runs in FP64 with the oracle's operation order (no FMA
// contraction: __dmul_rn/__dadd_rn, IEEE sqrt/div), so verdicts are bit-identical to the CPU path
// up to CUDA's sin/cos in the forward kinematics (<= 2 ulp).  FP32 is only used for conservative
// culling: an OBB-OBB separating-axis prefilter in front of each MPR pair and the environment
// hierarchy traversal, both with a slack far above the FP32 error, so they can only skip tests
// whose answer is "disjoint".
//
// Four kernels per batch, cheap FP32 culling and exact FP64 tests kept apart so that warps stay convergent:
//   1 fetchFkKernel   (thread per state)            FK -> 12 frames, floor clearance, box culling of the 24 pairs
//   2 fetchPairKernel (thread per surviving pair)   exact shape-shape test (MPR / box-box)
//   3 fetchEnvKernel  (thread per state x shape)    FP32 traversal of the environment hierarchy
//   4 fetchLeafKernel (thread per surviving leaf)   exact shape-triangle MPR
// Survivors travel through device-side work lists; a full list degrades to inline exact tests.
#include "fetch_device.cuh"
#include <cstdlib>
#include "device_math.cuh"

#include <algorithm>
#include <cfloat>

namespace mplb {
namespace {

struct FrameD { double R[3][3]; double t[3]; };

// ---- Eigen-order frame algebra (oracle/mplb_oracle_fetch.inc: frame_translate / frame_rotate / frame_mul)
__device__ void frameTranslate(FrameD &f, double x, double y, double z) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
        f.t[i] = dadd(f.t[i], dadd(dadd(dmul(f.R[i][0], x), dmul(f.R[i][1], y)), dmul(f.R[i][2], z)));
}
__device__ void mat3Mul(const double A[3][3], const double B[3][3], double C[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i][j] = dadd(dadd(dmul(A[i][0], B[0][j]), dmul(A[i][1], B[1][j])), dmul(A[i][2], B[2][j]));
}
// Eigen AngleAxis::toRotationMatrix about a unit coordinate axis, then linear *= M
__device__ void frameRotate(FrameD &f, double angle, int axis) {
    double ax[3] = {0, 0, 0};
    ax[axis] = 1;
    double s, c;
    sincos(angle, &s, &c);
    const double sa[3] = {dmul(s, ax[0]), dmul(s, ax[1]), dmul(s, ax[2])};
    const double omc = dsub(1, c);
    const double c1[3] = {dmul(omc, ax[0]), dmul(omc, ax[1]), dmul(omc, ax[2])};
    double M[3][3], tmp;
    tmp = dmul(c1[0], ax[1]); M[0][1] = dsub(tmp, sa[2]); M[1][0] = dadd(tmp, sa[2]);
    tmp = dmul(c1[0], ax[2]); M[0][2] = dadd(tmp, sa[1]); M[2][0] = dsub(tmp, sa[1]);
    tmp = dmul(c1[1], ax[2]); M[1][2] = dsub(tmp, sa[0]); M[2][1] = dadd(tmp, sa[0]);
#pragma unroll
    for (int i = 0; i < 3; ++i) M[i][i] = dadd(dmul(c1[i], ax[i]), c);
    double L[3][3];
    mat3Mul(f.R, M, L);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) f.R[i][j] = L[i][j];
}
__device__ void frameMul(const FrameD &a, const double *b /* row-major 3x4 */, FrameD &r) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
            r.R[i][j] = dadd(dadd(dmul(a.R[i][0], b[0 + j]), dmul(a.R[i][1], b[4 + j])), dmul(a.R[i][2], b[8 + j]));
        r.t[i] = dadd(dadd(dadd(dmul(a.R[i][0], b[3]), dmul(a.R[i][1], b[7])), dmul(a.R[i][2], b[11])), a.t[i]);
    }
}

// ---- libccd objects ([EXT] fcl gjk_libccd-inl.h), mirror of the oracle's ccd_obj
struct CcdObj {
    int kind;  // 0 box, 1 capsule, 2 cylinder, 3 triangle
    double pos[3], rot[4], rotInv[4];
    double a, b, c;
    D3 p[3], cen;
};

constexpr double kCcdEps = DBL_EPSILON;
__device__ __forceinline__ bool ccdIsZero(double x) { return fabs(x) < kCcdEps; }
__device__ __forceinline__ bool ccdEq(double a, double b) {
    const double ab = fabs(dsub(a, b));
    if (ab < kCcdEps) return true;
    a = fabs(a);
    b = fabs(b);
    return b > a ? ab < dmul(kCcdEps, b) : ab < dmul(kCcdEps, a);
}
__device__ __forceinline__ bool ccdVecEq(D3 a, D3 b) { return ccdEq(a.x, b.x) && ccdEq(a.y, b.y) && ccdEq(a.z, b.z); }
__device__ __forceinline__ double ccdSign(double x) { return ccdIsZero(x) ? 0.0 : (x < 0 ? -1.0 : 1.0); }
__device__ __forceinline__ D3 d3add(D3 a, D3 b) { return {dadd(a.x, b.x), dadd(a.y, b.y), dadd(a.z, b.z)}; }
__device__ __forceinline__ D3 ccdNormalize(D3 v) {
    const double d = __ddiv_rn(1.0, __dsqrt_rn(d3dot(v, v)));
    return {dmul(v.x, d), dmul(v.y, d), dmul(v.z, d)};
}
__device__ __forceinline__ D3 d3neg(D3 v) { return {dmul(v.x, -1.0), dmul(v.y, -1.0), dmul(v.z, -1.0)}; }
// ccdQuatRotVec (libccd 2.1)
__device__ D3 quatRot(D3 v, const double q[4]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double c1x = dadd(dsub(dmul(y, v.z), dmul(z, v.y)), dmul(w, v.x));
    const double c1y = dadd(dsub(dmul(z, v.x), dmul(x, v.z)), dmul(w, v.y));
    const double c1z = dadd(dsub(dmul(x, v.y), dmul(y, v.x)), dmul(w, v.z));
    const double c2x = dsub(dmul(y, c1z), dmul(z, c1y));
    const double c2y = dsub(dmul(z, c1x), dmul(x, c1z));
    const double c2z = dsub(dmul(x, c1y), dmul(y, c1x));
    return {dadd(v.x, dmul(2, c2x)), dadd(v.y, dmul(2, c2y)), dadd(v.z, dmul(2, c2z))};
}
// Eigen Quaternion(Matrix3) [EXT]
__device__ void quatFromR(const double m[3][3], double q[4]) {
    double t = dadd(dadd(m[0][0], m[1][1]), m[2][2]);
    if (t > 0) {
        t = __dsqrt_rn(dadd(t, 1.0));
        q[3] = dmul(0.5, t);
        t = __ddiv_rn(0.5, t);
        q[0] = dmul(dsub(m[2][1], m[1][2]), t);
        q[1] = dmul(dsub(m[0][2], m[2][0]), t);
        q[2] = dmul(dsub(m[1][0], m[0][1]), t);
    } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = __dsqrt_rn(dadd(dsub(dsub(m[i][i], m[j][j]), m[k][k]), 1.0));
        q[i] = dmul(0.5, t);
        t = __ddiv_rn(0.5, t);
        q[3] = dmul(dsub(m[k][j], m[j][k]), t);
        q[j] = dmul(dadd(m[j][i], m[i][j]), t);
        q[k] = dmul(dadd(m[k][i], m[i][k]), t);
    }
}
__device__ void objSetFrame(CcdObj &o, const FrameD &f) {
    o.pos[0] = f.t[0]; o.pos[1] = f.t[1]; o.pos[2] = f.t[2];
    quatFromR(f.R, o.rot);
    const double len2 = dadd(dadd(dadd(dmul(o.rot[0], o.rot[0]), dmul(o.rot[1], o.rot[1])), dmul(o.rot[2], o.rot[2])),
                             dmul(o.rot[3], o.rot[3]));
    const double s = __ddiv_rn(1.0, len2);
    o.rotInv[0] = dmul(-o.rot[0], s);
    o.rotInv[1] = dmul(-o.rot[1], s);
    o.rotInv[2] = dmul(-o.rot[2], s);
    o.rotInv[3] = dmul(o.rot[3], s);
}
__device__ D3 objSupport(const CcdObj &o, D3 dir_) {
    const D3 dir = quatRot(dir_, o.rotInv);
    D3 v;
    if (o.kind == 0) {
        v = {dmul(ccdSign(dir.x), o.a), dmul(ccdSign(dir.y), o.b), dmul(ccdSign(dir.z), o.c)};
    } else if (o.kind == 1) {
        D3 n = ccdNormalize(dir);
        n = {dmul(n.x, o.a), dmul(n.y, o.a), dmul(n.z, o.a)};
        v = dir.z > 0 ? D3{dadd(0, n.x), dadd(0, n.y), dadd(o.b, n.z)} : D3{dadd(0, n.x), dadd(0, n.y), dadd(-o.b, n.z)};
    } else if (o.kind == 2) {
        const double zdist = __dsqrt_rn(dadd(dmul(dir.x, dir.x), dmul(dir.y, dir.y)));
        if (ccdIsZero(zdist)) {
            v = {0, 0, dmul(ccdSign(dir.z), o.b)};
        } else {
            const double rad = __ddiv_rn(o.a, zdist);
            v = {dmul(rad, dir.x), dmul(rad, dir.y), dmul(ccdSign(dir.z), o.b)};
        }
    } else {
        double maxdot = -DBL_MAX;
        v = o.p[0];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const D3 p = {dsub(o.p[i].x, o.cen.x), dsub(o.p[i].y, o.cen.y), dsub(o.p[i].z, o.cen.z)};
            const double d = d3dot(dir, p);
            if (d > maxdot) {
                v = o.p[i];
                maxdot = d;
            }
        }
    }
    v = quatRot(v, o.rot);
    return {dadd(v.x, o.pos[0]), dadd(v.y, o.pos[1]), dadd(v.z, o.pos[2])};
}
__device__ D3 objCenter(const CcdObj &o) {
    if (o.kind != 3) return {o.pos[0], o.pos[1], o.pos[2]};
    const D3 c = quatRot(o.cen, o.rot);
    return {dadd(c.x, o.pos[0]), dadd(c.y, o.pos[1]), dadd(c.z, o.pos[2])};
}
__device__ D3 minkSupport(const CcdObj &o1, const CcdObj &o2, D3 dir) {
    const D3 a = objSupport(o1, dir);
    const D3 b = objSupport(o2, d3neg(dir));
    return d3sub(a, b);
}

// ---- [EXT] libccd ccdMPRIntersect (discoverPortal + refinePortal), statement for statement as the oracle
__device__ int mprDiscoverPortal(const CcdObj &o1, const CcdObj &o2, D3 P[4]) {
    const D3 origin = {0, 0, 0};
    D3 dir, va, vb;
    double dot;
    P[0] = d3sub(objCenter(o1), objCenter(o2));
    if (ccdVecEq(P[0], origin)) P[0] = d3add(P[0], D3{dmul(kCcdEps, 10.0), 0, 0});
    dir = ccdNormalize(d3neg(P[0]));
    P[1] = minkSupport(o1, o2, dir);
    dot = d3dot(P[1], dir);
    if (ccdIsZero(dot) || dot < 0) return -1;
    dir = d3cross(P[0], P[1]);
    if (ccdIsZero(d3dot(dir, dir))) return ccdVecEq(P[1], origin) ? 1 : 2;
    dir = ccdNormalize(dir);
    P[2] = minkSupport(o1, o2, dir);
    dot = d3dot(P[2], dir);
    if (ccdIsZero(dot) || dot < 0) return -1;
    va = d3sub(P[1], P[0]);
    vb = d3sub(P[2], P[0]);
    dir = ccdNormalize(d3cross(va, vb));
    dot = d3dot(dir, P[0]);
    if (dot > 0) {
        const D3 t = P[1];
        P[1] = P[2];
        P[2] = t;
        dir = d3neg(dir);
    }
    for (int guard = 0; guard < 10000; ++guard) {
        bool cont = false;
        P[3] = minkSupport(o1, o2, dir);
        dot = d3dot(P[3], dir);
        if (ccdIsZero(dot) || dot < 0) return -1;
        va = d3cross(P[1], P[3]);
        dot = d3dot(va, P[0]);
        if (dot < 0 && !ccdIsZero(dot)) {
            P[2] = P[3];
            cont = true;
        }
        if (!cont) {
            va = d3cross(P[3], P[2]);
            dot = d3dot(va, P[0]);
            if (dot < 0 && !ccdIsZero(dot)) {
                P[1] = P[3];
                cont = true;
            }
        }
        if (!cont) return 0;
        va = d3sub(P[1], P[0]);
        vb = d3sub(P[2], P[0]);
        dir = ccdNormalize(d3cross(va, vb));
    }
    return -2;
}
__device__ int mprRefinePortal(const CcdObj &o1, const CcdObj &o2, D3 P[4], double tol) {
    for (int guard = 0; guard < 10000; ++guard) {
        const D3 dir = ccdNormalize(d3cross(d3sub(P[2], P[1]), d3sub(P[3], P[1])));
        double dot = d3dot(dir, P[1]);
        if (ccdIsZero(dot) || dot > 0) return 0;
        const D3 v4 = minkSupport(o1, o2, dir);
        dot = d3dot(v4, dir);
        if (!(ccdIsZero(dot) || dot > 0)) return -1;
        {
            const double dv1 = d3dot(P[1], dir), dv2 = d3dot(P[2], dir), dv3 = d3dot(P[3], dir), dv4 = d3dot(v4, dir);
            double d1 = dsub(dv4, dv1);
            const double d2 = dsub(dv4, dv2), d3 = dsub(dv4, dv3);
            d1 = fmin(d1, d2);
            d1 = fmin(d1, d3);
            if (ccdEq(d1, tol) || d1 < tol) return -1;
        }
        {
            const D3 v4v0 = d3cross(v4, P[0]);
            dot = d3dot(P[1], v4v0);
            if (dot > 0) {
                dot = d3dot(P[2], v4v0);
                if (dot > 0) P[1] = v4; else P[3] = v4;
            } else {
                dot = d3dot(P[3], v4v0);
                if (dot > 0) P[2] = v4; else P[1] = v4;
            }
        }
    }
    return -2;
}
// 1 intersect, 0 disjoint, -1 iteration guard tripped (reported, never silent)
__device__ __noinline__ int mprIntersect(const CcdObj &o1, const CcdObj &o2) {
    D3 P[4];
    const int res = mprDiscoverPortal(o1, o2, P);
    if (res == -2) return -1;
    if (res < 0) return 0;
    if (res > 0) return 1;
    const int r = mprRefinePortal(o1, o2, P, 1e-6);
    if (r == -2) return -1;
    return r == 0 ? 1 : 0;
}

// ---- the same algorithm, resumable: one support evaluation per call ---------------------------------------------------
// ccdMPRIntersect is a chain of Minkowski support evaluations (two shape supports each, by far the bulk of the
// arithmetic) with a few cross and dot products in between that decide whether it goes on.  Run to completion per lane,
// a warp serialises over the lanes' different iteration counts (ncu: 8.7 of 32 lanes).  MprState holds what survives
// between two support evaluations -- the portal, the next direction and where in the algorithm the lane is -- so that a
// warp can evaluate one support for all of its lanes together, whatever phase each is in, and hand a lane the next work
// item the moment its test is decided (fetchMprKernel).  Every operation and its order are those of mprDiscoverPortal /
// mprRefinePortal above: same results, bit for bit.
struct MprState {
    D3 P[4];
    D3 dir;
    int phase;   // 0 idle, 1..3 discoverPortal's three stages, 4 refinePortal
    int guard;
};
enum : int { kMprIdle = 0, kMprD1 = 1, kMprD2 = 2, kMprD3 = 3, kMprR = 4 };

__device__ __forceinline__ void mprBegin(const CcdObj &o1, const CcdObj &o2, MprState &m) {
    const D3 origin = {0, 0, 0};
    m.P[0] = d3sub(objCenter(o1), objCenter(o2));
    if (ccdVecEq(m.P[0], origin)) m.P[0] = d3add(m.P[0], D3{dmul(kCcdEps, 10.0), 0, 0});
    m.dir = ccdNormalize(d3neg(m.P[0]));
    m.phase = kMprD1;
    m.guard = 0;
}
// refinePortal's loop head: the portal's direction, and whether the origin already lies inside -> true: intersecting
__device__ __forceinline__ bool mprRefineHead(MprState &m) {
    m.dir = ccdNormalize(d3cross(d3sub(m.P[2], m.P[1]), d3sub(m.P[3], m.P[1])));
    const double dot = d3dot(m.dir, m.P[1]);
    return ccdIsZero(dot) || dot > 0;
}
// S = support of the Minkowski difference in m.dir.  -> -2: goes on (m.dir is the next direction), 1 intersect,
// 0 disjoint, -1 iteration guard tripped
__device__ __forceinline__ int mprAdvance(MprState &m, const D3 S) {
    double dot;
    D3 va, vb;
    if (m.phase == kMprD1) {
        m.P[1] = S;
        dot = d3dot(m.P[1], m.dir);
        if (ccdIsZero(dot) || dot < 0) return 0;
        m.dir = d3cross(m.P[0], m.P[1]);
        if (ccdIsZero(d3dot(m.dir, m.dir))) return 1;   // (discoverPortal's 1 and 2 both mean "intersecting")
        m.dir = ccdNormalize(m.dir);
        m.phase = kMprD2;
        return -2;
    }
    if (m.phase == kMprD2) {
        m.P[2] = S;
        dot = d3dot(m.P[2], m.dir);
        if (ccdIsZero(dot) || dot < 0) return 0;
        va = d3sub(m.P[1], m.P[0]);
        vb = d3sub(m.P[2], m.P[0]);
        m.dir = ccdNormalize(d3cross(va, vb));
        dot = d3dot(m.dir, m.P[0]);
        if (dot > 0) {
            const D3 t = m.P[1];
            m.P[1] = m.P[2];
            m.P[2] = t;
            m.dir = d3neg(m.dir);
        }
        m.phase = kMprD3;
        m.guard = 0;
        return -2;
    }
    if (m.phase == kMprD3) {
        bool cont = false;
        m.P[3] = S;
        dot = d3dot(m.P[3], m.dir);
        if (ccdIsZero(dot) || dot < 0) return 0;
        va = d3cross(m.P[1], m.P[3]);
        dot = d3dot(va, m.P[0]);
        if (dot < 0 && !ccdIsZero(dot)) {
            m.P[2] = m.P[3];
            cont = true;
        }
        if (!cont) {
            va = d3cross(m.P[3], m.P[2]);
            dot = d3dot(va, m.P[0]);
            if (dot < 0 && !ccdIsZero(dot)) {
                m.P[1] = m.P[3];
                cont = true;
            }
        }
        if (!cont) {   // portal found: on to refinePortal
            m.phase = kMprR;
            m.guard = 0;
            return mprRefineHead(m) ? 1 : -2;
        }
        if (++m.guard >= 10000) return -1;
        va = d3sub(m.P[1], m.P[0]);
        vb = d3sub(m.P[2], m.P[0]);
        m.dir = ccdNormalize(d3cross(va, vb));
        return -2;
    }
    // refinePortal, after its support
    const D3 v4 = S;
    dot = d3dot(v4, m.dir);
    if (!(ccdIsZero(dot) || dot > 0)) return 0;
    {
        const double dv1 = d3dot(m.P[1], m.dir), dv2 = d3dot(m.P[2], m.dir), dv3 = d3dot(m.P[3], m.dir), dv4 = d3dot(v4, m.dir);
        double d1 = dsub(dv4, dv1);
        const double d2 = dsub(dv4, dv2), d3 = dsub(dv4, dv3);
        d1 = fmin(d1, d2);
        d1 = fmin(d1, d3);
        if (ccdEq(d1, 1e-6) || d1 < 1e-6) return 0;
    }
    {
        const D3 v4v0 = d3cross(v4, m.P[0]);
        dot = d3dot(m.P[1], v4v0);
        if (dot > 0) {
            dot = d3dot(m.P[2], v4v0);
            if (dot > 0) m.P[1] = v4; else m.P[3] = v4;
        } else {
            dot = d3dot(m.P[3], v4v0);
            if (dot > 0) m.P[2] = v4; else m.P[1] = v4;
        }
    }
    if (++m.guard >= 10000) return -1;
    return mprRefineHead(m) ? 1 : -2;
}

// ---- [EXT] fcl boxBox2 reduced to its boolean result
__device__ bool boxBoxIntersect(const FetchShape &s1, const FrameD &f1, const FetchShape &s2, const FrameD &f2) {
    const double p[3] = {dsub(f2.t[0], f1.t[0]), dsub(f2.t[1], f1.t[1]), dsub(f2.t[2], f1.t[2])};
    double pp[3], R[3][3], Q[3][3];
    const double A[3] = {s1.a, s1.b, s1.c}, B[3] = {s2.a, s2.b, s2.c};
    for (int i = 0; i < 3; ++i) pp[i] = dadd(dadd(dmul(f1.R[0][i], p[0]), dmul(f1.R[1][i], p[1])), dmul(f1.R[2][i], p[2]));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            R[i][j] = dadd(dadd(dmul(f1.R[0][i], f2.R[0][j]), dmul(f1.R[1][i], f2.R[1][j])), dmul(f1.R[2][i], f2.R[2][j]));
            Q[i][j] = fabs(R[i][j]);
        }
    for (int i = 0; i < 3; ++i)
        if (dsub(fabs(pp[i]), dadd(dadd(dadd(dmul(Q[i][0], B[0]), dmul(Q[i][1], B[1])), dmul(Q[i][2], B[2])), A[i])) > 0)
            return false;
    for (int j = 0; j < 3; ++j) {
        const double tmp = dadd(dadd(dmul(f2.R[0][j], p[0]), dmul(f2.R[1][j], p[1])), dmul(f2.R[2][j], p[2]));
        if (dsub(fabs(tmp), dadd(dadd(dadd(dmul(Q[0][j], A[0]), dmul(Q[1][j], A[1])), dmul(Q[2][j], A[2])), B[j])) > 0)
            return false;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Q[i][j] = dadd(Q[i][j], 1e-6);
    for (int i = 0; i < 3; ++i) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        for (int j = 0; j < 3; ++j) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            const double tmp = dsub(dmul(pp[i2], R[i1][j]), dmul(pp[i1], R[i2][j]));
            const double s2 = dsub(fabs(tmp), dadd(dadd(dadd(dmul(A[i1], Q[i2][j]), dmul(A[i2], Q[i1][j])), dmul(B[j1], Q[i][j2])),
                                                   dmul(B[j2], Q[i][j1])));
            if (s2 > kCcdEps) return false;
        }
    }
    return true;
}

// ---- conservative FP32 culling helpers -------------------------------------------------------------
// half extents of the shape's oriented box ([EXT] fcl computeBV for Box / Capsule / Cylinder)
__device__ __forceinline__ void shapeExtents(const FetchShape &s, float e[3]) {
    if (s.type == 0) { e[0] = (float)s.a; e[1] = (float)s.b; e[2] = (float)s.c; }
    else if (s.type == 1) { e[0] = (float)s.a; e[1] = (float)s.a; e[2] = (float)(s.b + s.a); }
    else { e[0] = (float)s.a; e[1] = (float)s.a; e[2] = (float)s.b; }
}
constexpr float kCullSlack = 1e-4f;  // metres; FP32 error of these tests is ~1e-6
#ifndef MPLB_FETCH_TRI_SAT
#define MPLB_FETCH_TRI_SAT 1
#endif

// Separating-axis test of a shape's oriented box (half extents e, centred at the origin of its frame) against a triangle
// whose vertices q[0..2] are given in that frame: the box's three axes, the triangle's normal, the nine edge cross
// products.  true only when the two are farther apart than kCullSlack along one of them (|axis|_1 >= |axis|_2 scales the
// slack; a vanishing cross product separates nothing), so -- the shape lying inside its box -- the exact test would
// answer "disjoint": the same argument, slack and arithmetic as the box tests of the traversal.
__device__ __forceinline__ bool boxTriangleApart(const float e[3], const float q[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float mn = fminf(q[0][i], fminf(q[1][i], q[2][i])), mx = fmaxf(q[0][i], fmaxf(q[1][i], q[2][i]));
        if (mn > e[i] + kCullSlack || mx < -e[i] - kCullSlack) return true;
    }
    float ed[3][3];   // edges q1 - q0, q2 - q1, q0 - q2
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ed[0][i] = q[1][i] - q[0][i];
        ed[1][i] = q[2][i] - q[1][i];
        ed[2][i] = q[0][i] - q[2][i];
    }
    {
        const float n[3] = {ed[0][1] * ed[1][2] - ed[0][2] * ed[1][1], ed[0][2] * ed[1][0] - ed[0][0] * ed[1][2],
                            ed[0][0] * ed[1][1] - ed[0][1] * ed[1][0]};
        const float d = n[0] * q[0][0] + n[1] * q[0][1] + n[2] * q[0][2];
        const float a0 = fabsf(n[0]), a1 = fabsf(n[1]), a2 = fabsf(n[2]);
        if (fabsf(d) > e[0] * a0 + e[1] * a1 + e[2] * a2 + kCullSlack * (a0 + a1 + a2)) return true;
    }
    bool apart = false;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float fx = ed[j][0], fy = ed[j][1], fz = ed[j][2];
        const float L[3][3] = {{0.f, -fz, fy}, {fz, 0.f, -fx}, {-fy, fx, 0.f}};   // u_i x f for the three box axes
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float p0 = L[i][0] * q[0][0] + L[i][1] * q[0][1] + L[i][2] * q[0][2];
            const float p1 = L[i][0] * q[1][0] + L[i][1] * q[1][1] + L[i][2] * q[1][2];
            const float p2 = L[i][0] * q[2][0] + L[i][1] * q[2][1] + L[i][2] * q[2][2];
            const float a0 = fabsf(L[i][0]), a1 = fabsf(L[i][1]), a2 = fabsf(L[i][2]);
            const float r = e[0] * a0 + e[1] * a1 + e[2] * a2 + kCullSlack * (a0 + a1 + a2);
            apart |= fminf(p0, fminf(p1, p2)) > r || fmaxf(p0, fmaxf(p1, p2)) < -r;
        }
    }
    return apart;
}

// A sphere-swept segment that contains the shape: exact for a capsule, the cylinder with its rim rounded, a box
// swept along its longest side.  For the slender arm links this bounds distances far better than their boxes.
__device__ __forceinline__ void shapeSweep(const FetchShape &s, int &axis, float &half, float &radius) {
    if (s.type != 0) {
        axis = 2;
        half = (float)s.b;
        radius = (float)s.a * 1.00001f;
        return;
    }
    const float a = (float)s.a, b = (float)s.b, c = (float)s.c;
    axis = (a >= b && a >= c) ? 0 : (b >= c ? 1 : 2);
    half = axis == 0 ? a : axis == 1 ? b : c;
    const float o1 = axis == 0 ? b : a, o2 = axis == 2 ? b : c;
    radius = sqrtf(o1 * o1 + o2 * o2) * 1.00001f;
}

// lower bound of the distance between the segment p0 p1 and the box [-e, e], both in the box's frame: the distance
// from a point of the segment to the box is convex along the segment and 1-Lipschitz, so a ternary search brackets
// its minimum and the bracket's length bounds what the search can still be off by
// (cylinder == true: the solid is the cylinder of radius e[0] and half length e[2] around the frame's z axis)
__device__ __forceinline__ float segmentBoxDistance(const float p0[3], const float p1[3], const float e[3], bool cylinder = false) {
    auto at = [&](float t) {
        const float x = p0[0] + t * (p1[0] - p0[0]), y = p0[1] + t * (p1[1] - p0[1]), z = p0[2] + t * (p1[2] - p0[2]);
        if (cylinder) {
            const float dr = fmaxf(sqrtf(x * x + y * y) - e[0], 0.f), dz = fmaxf(fabsf(z) - e[2], 0.f);
            return sqrtf(dr * dr + dz * dz);
        }
        const float wx = fmaxf(fabsf(x) - e[0], 0.f), wy = fmaxf(fabsf(y) - e[1], 0.f), wz = fmaxf(fabsf(z) - e[2], 0.f);
        return sqrtf(wx * wx + wy * wy + wz * wz);
    };
    float lo = 0.f, hi = 1.f;
#pragma unroll 1
    for (int it = 0; it < 14; ++it) {
        const float m1 = lo + (hi - lo) * (1.f / 3.f), m2 = hi - (hi - lo) * (1.f / 3.f);
        if (at(m1) <= at(m2)) hi = m2;
        else lo = m1;
    }
    const float len = sqrtf((p1[0] - p0[0]) * (p1[0] - p0[0]) + (p1[1] - p0[1]) * (p1[1] - p0[1]) + (p1[2] - p0[2]) * (p1[2] - p0[2]));
    return at(0.5f * (lo + hi)) - (hi - lo) * len;
}

// distance between the segments c1 +- h1 u1 and c2 +- h2 u2 (Ericson, Real-Time Collision Detection 5.1.9)
__device__ __forceinline__ float segmentDistance(const float c1[3], const float u1[3], float h1, const float c2[3],
                                                 const float u2[3], float h2) {
    float d1[3], d2[3], r[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        d1[i] = 2.f * h1 * u1[i];
        d2[i] = 2.f * h2 * u2[i];
        r[i] = (c1[i] - h1 * u1[i]) - (c2[i] - h2 * u2[i]);
    }
    const float a = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2], e = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2];
    const float f = d2[0] * r[0] + d2[1] * r[1] + d2[2] * r[2], c = d1[0] * r[0] + d1[1] * r[1] + d1[2] * r[2];
    const float b = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2];
    float sN = 0.f, tN = 0.f;
    const float eps = 1e-12f;
    if (a <= eps && e <= eps) {
        sN = tN = 0.f;
    } else if (a <= eps) {
        tN = fminf(fmaxf(f / e, 0.f), 1.f);
    } else {
        if (e <= eps) {
            sN = fminf(fmaxf(-c / a, 0.f), 1.f);
        } else {
            const float den = a * e - b * b;
            sN = den > eps * a * e ? fminf(fmaxf((b * f - c * e) / den, 0.f), 1.f) : 0.f;
            tN = (b * sN + f) / e;
            if (tN < 0.f) {
                tN = 0.f;
                sN = fminf(fmaxf(-c / a, 0.f), 1.f);
            } else if (tN > 1.f) {
                tN = 1.f;
                sN = fminf(fmaxf((b - c) / a, 0.f), 1.f);
            }
        }
    }
    float dd = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float w = r[i] + sN * d1[i] - tN * d2[i];
        dd += w * w;
    }
    return sqrtf(dd);
}

__device__ void makeShapeObj(const FetchShape &s, const FrameD &f, CcdObj &o) {
    o.kind = s.type;
    o.a = s.a; o.b = s.b; o.c = s.c;
    objSetFrame(o, f);
}

// fetch_robot.hpp:508-577 pair order (indices into the geometry list of fetch_robot.hpp:38-53)
__constant__ unsigned char kSelfPairs[24][2] = {
    {0, 6}, {0, 7}, {0, 8}, {0, 9}, {1, 6}, {1, 7}, {1, 8}, {1, 9}, {2, 6}, {2, 7}, {2, 8}, {2, 9},
    {3, 8}, {4, 8}, {4, 9}, {4, 11}, {5, 11}, {6, 11}, {7, 10}, {7, 11}, {8, 10}, {8, 11}, {9, 10}, {9, 11}};

// A frame is three 32-byte rows (R[i][0..2], t[i]), 32-byte aligned: one 256-bit access per row (sm_100) instead of
// four 64-bit ones -- every thread works on its own state, so each access is a separate L1 wavefront either way
__device__ void loadFrame(const double *src, FrameD &f) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(f.R[i][0]), "=d"(f.R[i][1]), "=d"(f.R[i][2]), "=d"(f.t[i])
                     : "l"(src + 4 * i));
}
__device__ void storeFrame(double *dst, const FrameD &f) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4 * i), "d"(f.R[i][0]), "d"(f.R[i][1]), "d"(f.R[i][2]),
                     "d"(f.t[i])
                     : "memory");
}

// Work lists filled by the culling kernels and drained by the exact (FP64 MPR) kernels: separating the
// cheap FP32 culling from the expensive exact tests keeps every warp of either kind convergent.
struct WorkList {
    unsigned long long *items;   // packed work descriptors
    unsigned *count;             // number of items appended (may exceed capacity: overflow handled inline)
    unsigned capacity;
};
__device__ __forceinline__ bool listPush(const WorkList &l, unsigned long long v) {
    const unsigned at = atomicAdd(l.count, 1u);
    if (at >= l.capacity) return false;
    l.items[at] = v;
    return true;
}

__device__ int exactPair(const FetchSceneDev &sc, const double *frames, unsigned long long state, int k) {
    const int a = kSelfPairs[k][0], b = kSelfPairs[k][1];
    const FetchShape &sa = sc.shapes[a], &sb = sc.shapes[b];
    FrameD fa, fb;
    loadFrame(frames + state * 144 + 12 * a, fa);
    loadFrame(frames + state * 144 + 12 * b, fb);
    if (sa.type == 0 && sb.type == 0) return boxBoxIntersect(sa, fa, sb, fb) ? 1 : 0;
    CcdObj o1, o2;
    makeShapeObj(sa, fa, o1);
    makeShapeObj(sb, fb, o2);
    return mprIntersect(o1, o2);
}

// the two libccd objects of a (state, geometry, triangle) item / of a (state, pair) item
// (the triangle's frame is the environment's: its quaternion and inverse are the same for every item -- envObjFrame
// forms them once per thread, leafObjects only fills in the triangle)
__device__ void envObjFrame(const FetchSceneDev &sc, CcdObj &t) {
    FrameD ef;
#pragma unroll
    for (int i = 0; i < 3; ++i) {   // (kernel-parameter space, not global memory)
#pragma unroll
        for (int j = 0; j < 3; ++j) ef.R[i][j] = sc.envFrame[4 * i + j];
        ef.t[i] = sc.envFrame[4 * i + 3];
    }
    t.kind = 3;
    t.a = t.b = t.c = 0;
    objSetFrame(t, ef);
}
__device__ void leafObjects(const FetchSceneDev &sc, const double *frames, unsigned long long state, int g, unsigned tri,
                            CcdObj &shapeObj, CcdObj &t) {
    FrameD f;
    loadFrame(frames + state * 144 + 12 * g, f);
    makeShapeObj(sc.shapes[g], f, shapeObj);
    const double *tp = sc.envTris64 + 9 * (size_t)tri;
    t.p[0] = {tp[0], tp[1], tp[2]};
    t.p[1] = {tp[3], tp[4], tp[5]};
    t.p[2] = {tp[6], tp[7], tp[8]};
    t.cen = {__ddiv_rn(dadd(dadd(tp[0], tp[3]), tp[6]), 3.0), __ddiv_rn(dadd(dadd(tp[1], tp[4]), tp[7]), 3.0),
             __ddiv_rn(dadd(dadd(tp[2], tp[5]), tp[8]), 3.0)};
}
__device__ void pairObjects(const FetchSceneDev &sc, const double *frames, unsigned long long state, int k, CcdObj &o1,
                            CcdObj &o2) {
    const int a = kSelfPairs[k][0], b = kSelfPairs[k][1];
    FrameD fa, fb;
    loadFrame(frames + state * 144 + 12 * a, fa);
    loadFrame(frames + state * 144 + 12 * b, fb);
    makeShapeObj(sc.shapes[a], fa, o1);
    makeShapeObj(sc.shapes[b], fb, o2);
}

__device__ int exactLeaf(const FetchSceneDev &sc, const double *frames, unsigned long long state, int g, unsigned tri) {
    FrameD f, ef;
    loadFrame(frames + state * 144 + 12 * g, f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {   // (kernel-parameter space, not global memory)
#pragma unroll
        for (int j = 0; j < 3; ++j) ef.R[i][j] = sc.envFrame[4 * i + j];
        ef.t[i] = sc.envFrame[4 * i + 3];
    }
    CcdObj shapeObj, t;
    makeShapeObj(sc.shapes[g], f, shapeObj);
    const double *tp = sc.envTris64 + 9 * (size_t)tri;
    t.kind = 3;
    t.a = t.b = t.c = 0;
    t.p[0] = {tp[0], tp[1], tp[2]};
    t.p[1] = {tp[3], tp[4], tp[5]};
    t.p[2] = {tp[6], tp[7], tp[8]};
    t.cen = {__ddiv_rn(dadd(dadd(tp[0], tp[3]), tp[6]), 3.0), __ddiv_rn(dadd(dadd(tp[1], tp[4]), tp[7]), 3.0),
             __ddiv_rn(dadd(dadd(tp[2], tp[5]), tp[8]), 3.0)};
    objSetFrame(t, ef);
    return mprIntersect(shapeObj, t);   // [EXT] GJKSolver_libccd::shapeTriangleIntersect
}

// Kernel 1, one thread per state: FK -> 12 frames, floor clearance, FP32 box culling of the 24 pairs.
// `states` are explicit joint vectors, or (edgeItems != nullptr) samples of edges: item = (edge, i)
// with state = from + (to - from) * (i * (1/steps)), i == steps meaning `to` itself.
__global__ void __launch_bounds__(128)
fetchFkKernel(FetchSceneDev sc, const double *__restrict__ states, const double *__restrict__ from,
              const double *__restrict__ to, const unsigned *__restrict__ steps,
              const unsigned long long *__restrict__ edgeItems, const float2 *__restrict__ slack,
              const unsigned *__restrict__ dead, unsigned long long n, double *__restrict__ frames /* [n][12][12] */,
              unsigned char *__restrict__ selfHit, unsigned char *__restrict__ touched, WorkList pairs, FetchCounters *ctr) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    // interval item: how far each geometry can stray from its pose at this sample (see FetchSceneDev::reach)
    const float2 sl = slack ? slack[i] : make_float2(0.f, 0.f);
    const bool interval = sl.x > 0.f || sl.y > 0.f;
    unsigned near = 0;   // why the interval is not certified: 1 floor, 2 pair within reach, 4 pair boxes overlap, 8 / 16 same for the environment
    double q[8];
    if (edgeItems) {
        const unsigned long long it = edgeItems[i];
        const unsigned e = (unsigned)(it >> 32), k = (unsigned)it;
        if (dead[e]) {
            selfHit[i] = 2;  // edge already decided: skip
            return;
        }
        const unsigned st = steps[e];
        if (k >= st) {
#pragma unroll
            for (int j = 0; j < 8; ++j) q[j] = to[8 * (size_t)e + j];
        } else {
            // interpolate.hpp:9-16: a + (b - a) * t with t = i * delta (fetch_scenario.hpp:146,151)
            const double t = dmul((double)k, __ddiv_rn(1.0, (double)st));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double a = from[8 * (size_t)e + j], b = to[8 * (size_t)e + j];
                q[j] = dadd(a, dmul(dsub(b, a), t));
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = states[8 * i + j];
    }

    // fk(): fetch_robot.hpp:276-315; every geometry frame goes straight to global memory
    double *out = frames + i * 144;
    FrameD cur, g;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int b = 0; b < 3; ++b) cur.R[a][b] = a == b ? 1.0 : 0.0;
        cur.t[a] = 0;
    }
    // conservative FP32 copy of each geometry box for the pair culling: centre + axes (columns of R)
    float bx[12][12];
    auto emit = [&](int gi) {
        frameMul(cur, sc.origins + 12 * gi, g);
        storeFrame(out + 12 * gi, g);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) bx[gi][3 * r + c] = (float)g.R[r][c];
            bx[gi][9 + r] = (float)g.t[r];
        }
    };
    emit(0);                                            // base
    frameTranslate(cur, -0.086875, 0, 0.37743);         // torsoLiftJointOrigin
    frameTranslate(cur, 0, 0, q[0]);                    // torsoLiftLink
    emit(1); emit(10); emit(11);                        // torso box, neck, head
    frameTranslate(cur, 0.119525, 0, 0.34858);
    frameRotate(cur, q[1], 2);                          // shoulderPanLink
    emit(2);
    frameTranslate(cur, 0.117, 0, 0.0599999999999999);
    frameRotate(cur, q[2], 1);                          // shoulderLiftLink
    emit(3);
    frameTranslate(cur, 0.219, 0, 0);
    frameRotate(cur, q[3], 0);                          // upperarmRollLink
    emit(4);
    frameTranslate(cur, 0.133, 0, 0);
    frameRotate(cur, q[4], 1);                          // elbowFlexLink
    emit(5);
    frameTranslate(cur, 0.197, 0, 0);
    frameRotate(cur, q[5], 0);                          // forearmRollLink
    emit(6);
    frameTranslate(cur, 0.1245, 0, 0);
    frameRotate(cur, q[6], 1);                          // wristFlexLink
    emit(7);
    frameTranslate(cur, 0.1385, 0, 0);
    frameRotate(cur, q[7], 0);                          // wristRollLink
    emit(8);
    frameTranslate(cur, 0.16645, 0, 0);                 // gripperAxis
    const double gripperAxisZ = cur.t[2];
    frameTranslate(cur, -0.09, 0, 0.0025);              // gripperLink == gripper geometry frame
    emit(9);                                            // origin 9 is the identity

    // selfCollision(): fetch_robot.hpp:502-584.  The floor test is exact here; pairs whose boxes are not
    // certainly disjoint go to the exact kernel (the verdict is an OR, so the order does not matter).
    int hit = gripperAxisZ < 0.30 ? 1 : 0;              // floorClearance_, fetch_robot.hpp:105,580
    // (the gripper axis point moves like the gripper geometry at most)
    if (interval && gripperAxisZ - 0.30 <= (double)(sl.x * sc.lift[9] + sl.y * sc.reach[9]) * 1.0001) near |= 1u;
    if (!hit) {
        for (int k = 0; k < 24; ++k) {
            const int a = kSelfPairs[k][0], b = kSelfPairs[k][1];
            float e1[3], e2[3];
            shapeExtents(sc.shapes[a], e1);
            shapeExtents(sc.shapes[b], e2);
            const float *A = bx[a], *B = bx[b];
            float R[3][3], T[3];
            const float d[3] = {B[9] - A[9], B[10] - A[10], B[11] - A[11]};
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) R[r][c] = A[r] * B[c] + A[3 + r] * B[3 + c] + A[6 + r] * B[6 + c];
                T[r] = A[r] * d[0] + A[3 + r] * d[1] + A[6 + r] * d[2];
            }
            float dist2;
            const float4 u0 = make_float4(1, 0, 0, 0), u1 = make_float4(0, 1, 0, 0), u2 = make_float4(0, 0, 1, 0);
            const float gap = obbGap(u0, u1, u2, make_float4(e1[0] * 1.00001f, e1[1] * 1.00001f, e1[2] * 1.00001f, 0), u0, u1, u2,
                                     make_float4(e2[0] * 1.00001f, e2[1] * 1.00001f, e2[2] * 1.00001f, 0),
                                     make_float4(R[0][0], R[0][1], R[0][2], T[0]), make_float4(R[1][0], R[1][1], R[1][2], T[1]),
                                     make_float4(R[2][0], R[2][1], R[2][2], T[2]), dist2);
            // how far the pair has to be apart for this item to be done with it (the interval's reach on top of the cull slack)
            const float need = kCullSlack + (interval ? (sl.x * (sc.lift[a] + sc.lift[b]) + sl.y * (sc.reach[a] + sc.reach[b])) * 1.0001f + 1e-6f : 0.f);
            // Boxes farther apart than that: nothing to do.  Otherwise the exact kernel looks closer (tighter distance
            // bounds first, MPR if those do not separate the shapes either) and reports back through `touched`.
            if (gap > need) continue;
            if (!listPush(pairs, (i << 8) | (unsigned)k)) {   // list full: decide it here
                near |= 4u;
                const int r = exactPair(sc, frames, i, k);
                if (r < 0) atomicExch(&ctr->error, 1u);
                if (r == 1) {
                    hit = 1;
                    break;
                }
            }
        }
    }
    selfHit[i] = (unsigned char)hit;
    if (touched && near) touched[i] = (unsigned char)near;
}

// Lower bound of the distance between geometries a and b of one state, from their FP32 frames (rotation row-major in
// [0..8], centre in [9..11]): the distance of the two swept segments, and -- when one shape is a box (torso, gripper)
// or a wide cylinder (base, head), whose swept segment is a poor bound (the head's adds a 25 cm dome above and below
// it) -- the distance from the other's segment to that solid itself.
__device__ float pairClearance(const FetchSceneDev &sc, int a, int b, const float *A, const float *B) {
    float e1[3], e2[3];
    shapeExtents(sc.shapes[a], e1);
    shapeExtents(sc.shapes[b], e2);
    int ax1, ax2;
    float h1, r1, h2, r2;
    shapeSweep(sc.shapes[a], ax1, h1, r1);
    shapeSweep(sc.shapes[b], ax2, h2, r2);
    const float u1[3] = {A[ax1], A[3 + ax1], A[6 + ax1]}, u2[3] = {B[ax2], B[3 + ax2], B[6 + ax2]};
    float clear = segmentDistance(A + 9, u1, h1, B + 9, u2, h2) * 0.99999f - (r1 + r2) - 2e-6f;
    const bool wideA = sc.shapes[a].type == 2 && sc.shapes[a].a > 0.1, wideB = sc.shapes[b].type == 2 && sc.shapes[b].a > 0.1;
    bool solidA = sc.shapes[a].type == 0, solidB = sc.shapes[b].type == 0;   // "use this shape as the solid"
    bool cyl = false;
    if (!solidA && !solidB && wideA != wideB) {
        solidA = wideA;
        solidB = wideB;
        cyl = true;
    } else if (solidA != solidB && (solidA ? wideB : wideA)) {   // box against wide cylinder: the cylinder is the solid
        solidA = wideA;
        solidB = wideB;
        cyl = true;
    }
    if (solidA != solidB) {
        const float *X = solidA ? A : B, *Y = solidA ? B : A;           // X: the solid, Y: the swept segment
        const float *uy = solidA ? u2 : u1;
        const float hy = solidA ? h2 : h1, ry = solidA ? r2 : r1;
        float p0[3], p1[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {   // X^T (y - cx): the solid's axes are the columns of X's rotation
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float dc = Y[9 + c] - X[9 + c];
                s0 += X[3 * c + r] * (dc - hy * uy[c]);
                s1 += X[3 * c + r] * (dc + hy * uy[c]);
            }
            p0[r] = s0;
            p1[r] = s1;
        }
        const float *ex = solidA ? e1 : e2;   // (a cylinder's extents are radius, radius, half length)
        clear = fmaxf(clear, segmentBoxDistance(p0, p1, ex, cyl) * 0.99999f - ry - 4e-6f);
    }
    return clear;
}

// Kernel 2, one thread per listed (state, pair) whose boxes are within reach of each other: tighter distance bounds
// first (they settle most of them), the exact shape-shape test if the shapes may touch at this very state
__global__ void __launch_bounds__(128)
fetchPairKernel(FetchSceneDev sc, const double *__restrict__ frames, WorkList pairs, WorkList exact,
                const float2 *__restrict__ slack, unsigned char *touched, unsigned char *selfHit, FetchCounters *ctr) {
    const unsigned n = min(*pairs.count, pairs.capacity);
    unsigned tested = 0;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const unsigned long long v = pairs.items[j];
        const unsigned long long state = v >> 8;
        if (selfHit[state]) continue;   // already colliding (floor, or another pair)
        const int k = (int)(v & 255), a = kSelfPairs[k][0], b = kSelfPairs[k][1];
        float A[12], B[12];
        {
            FrameD fa, fb;
            loadFrame(frames + state * 144 + 12 * a, fa);
            loadFrame(frames + state * 144 + 12 * b, fb);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    A[3 * r + c] = (float)fa.R[r][c];
                    B[3 * r + c] = (float)fb.R[r][c];
                }
                A[9 + r] = (float)fa.t[r];
                B[9 + r] = (float)fb.t[r];
            }
        }
        float need = kCullSlack;
        if (slack) {
            const float2 sl = slack[state];
            need += (sl.x * (sc.lift[a] + sc.lift[b]) + sl.y * (sc.reach[a] + sc.reach[b])) * 1.0001f + 1e-6f;
        }
        const float clear = pairClearance(sc, a, b, A, B);
        if (clear > need) continue;                  // apart, and over the whole interval
        if (clear > kCullSlack) {                    // apart here, but not certainly over the interval
            if (touched) touched[state] |= 2;
            continue;
        }
        if (touched) touched[state] |= 4;
        // MPR pairs go on to fetchMprKernel (a lane there never waits for another lane's longer test); the one Box x Box
        // pair -- and whatever does not fit the list -- is decided here
        if (!(sc.shapes[a].type == 0 && sc.shapes[b].type == 0) && listPush(exact, v)) continue;
        const int r = exactPair(sc, frames, state, k);
        ++tested;
        if (r < 0) atomicExch(&ctr->error, 1u);
        if (r == 1) selfHit[state] = 1;
    }
    if (tested) atomicAdd(&ctr->pairTests, (unsigned long long)tested);
}

// Kernels 2b / 4, persistent: the exact shape-shape (kLeaf == false: items (state << 8) | pair) and shape-triangle
// (kLeaf == true: items (state << 28) | geometry << 24 | triangle) tests as resumable MPR.  Every trip through the loop
// evaluates one Minkowski support for all the lanes of the warp; a lane whose test is decided takes the next item of the
// list at once.  hit: selfHit (bytes) / envHit (words) per state.
template <bool kLeaf>
__global__ void __launch_bounds__(128)
fetchMprKernel(FetchSceneDev sc, const double *__restrict__ frames, WorkList list, unsigned *cursor, unsigned char *selfHit,
               unsigned *envHit, FetchCounters *ctr) {
    const unsigned n = min(*list.count, list.capacity);
    const unsigned lane = threadIdx.x & 31u;
    CcdObj o1, o2;
    MprState m;
    m.phase = kMprIdle;
    if (kLeaf) envObjFrame(sc, o2);
    unsigned long long state = 0;
    unsigned tested = 0, chunkBase = 0, chunkLeft = 0, chunkStart = 0;
    unsigned long long chunkItem = 0;   // item chunkStart + lane
    unsigned nextBase = 0;              // (lane 0) the chunk claimed ahead
    bool nextClaimed = false;
    bool exhausted = false;
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, m.phase == kMprIdle);
        // (idle lanes waiting until several of them can set their tests up together made no difference: thresholds of
        // 2 .. 16 lanes all measured within the noise of refilling at once)
        if (need && !exhausted) {
            // Items are claimed 32 per warp -- the next chunk while the current one still has items, so that the cursor's
            // atomic is not waited for -- read one per lane, and what a lane will load when it is dealt the item (the
            // shape's frame, the triangle) is prefetched: the refill of a single lane used to be a chain of three
            // dependent global loads with the other lanes of the warp waiting (ncu: a fifth of this kernel's samples).
            if (chunkLeft == 0) {
                unsigned base = 0;
                if (lane == 0) base = nextClaimed ? nextBase : atomicAdd(cursor, 32u);
                nextClaimed = false;
                chunkBase = chunkStart = __shfl_sync(0xffffffffu, base, 0);
                chunkLeft = chunkBase < n ? min(32u, n - chunkBase) : 0u;
                if (chunkLeft == 0) exhausted = true;
                if (lane < chunkLeft) {
                    chunkItem = list.items[chunkBase + lane];
                    const unsigned long long st = kLeaf ? chunkItem >> 28 : chunkItem >> 8;
                    const double *fp = frames + st * 144;
                    if (kLeaf) {
                        fp += 12 * (int)((chunkItem >> 24) & 15);
                        const double *tp = sc.envTris64 + 9 * (size_t)(chunkItem & 0xffffff);
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(tp));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + 8));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(fp));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(fp + 11));
                    } else {
                        const int k = (int)(chunkItem & 255);
                        const double *fa = fp + 12 * kSelfPairs[k][0], *fb = fp + 12 * kSelfPairs[k][1];
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(fa));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(fa + 11));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(fb));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(fb + 11));
                    }
                }
            }
            const unsigned rank = __popc(need & ((1u << lane) - 1u));
            const unsigned take = min((unsigned)__popc(need), chunkLeft);
            const unsigned at = chunkBase + rank;
            chunkBase += take;
            chunkLeft -= take;
            const unsigned long long mine = __shfl_sync(0xffffffffu, chunkItem, (at - chunkStart) & 31u);
            if (chunkLeft < 12 && !nextClaimed && !exhausted) {
                if (lane == 0) nextBase = atomicAdd(cursor, 32u);
                nextClaimed = true;
            }
            if (m.phase == kMprIdle && rank < take) {
                const unsigned long long v = mine;
                state = kLeaf ? v >> 28 : v >> 8;
                const bool decided = kLeaf ? envHit[state] != 0u : selfHit[state] != 0;   // by another item of the state
                if (!decided) {
                    if (kLeaf) leafObjects(sc, frames, state, (int)((v >> 24) & 15), (unsigned)(v & 0xffffff), o1, o2);
                    else pairObjects(sc, frames, state, (int)(v & 255), o1, o2);
                    mprBegin(o1, o2, m);
                    ++tested;
                }
            }
        }
        const bool active = m.phase != kMprIdle;
        if (__ballot_sync(0xffffffffu, active) == 0) {
            if (exhausted) break;
            continue;
        }
        if (active) {
            const int r = mprAdvance(m, minkSupport(o1, o2, m.dir));
            if (r != -2) {
                if (r < 0) atomicExch(&ctr->error, 1u);
                if (r == 1) {
                    if (kLeaf) envHit[state] = 1u;
                    else selfHit[state] = 1;
                }
                m.phase = kMprIdle;
            }
        }
    }
    if (tested) atomicAdd(kLeaf ? &ctr->leafTests : &ctr->pairTests, (unsigned long long)tested);
}


// Kernel 3a: list the states that survived the self-collision tests (about half of a random batch)
__global__ void fetchSurvivorsKernel(const unsigned char *__restrict__ selfHit, unsigned long long n, unsigned *list,
                                     unsigned *count) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const bool keep = i < n && selfHit[i] == 0;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (!m) return;
    const unsigned lane = threadIdx.x & 31u;
    unsigned base = 0;
    if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (keep) list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned)i;
}

// Kernel 3b, persistent: every lane runs one depth-first traversal of the environment hierarchy (FP32 box
// tests against one shape's box) at a time and takes the next (surviving state, geometry) item the moment
// its stack runs dry, so lanes stay busy although traversal depths differ wildly.  Leaves that cannot be
// culled go to the exact kernel.
// kSmemNodes: the whole environment hierarchy is staged in shared memory (80-byte node stride: conflict free for
// lanes on different nodes).  The traversals are a dozen box tests long and were bound by the latency of the node
// loads (ncu: long-scoreboard 6.8 warps per issue at 10 warps / SM), which shared memory cuts by an order of
// magnitude; AUTOLAB's 1 256 nodes take 100 KB.  Larger environments use the global-memory variant.
// The kernel only culls: a leaf that cannot be culled goes to the exact kernel's list.  Should the list ever be full (it
// holds 16 leaf pairs per state on average; a state lists ~1.3) the launch group is flagged (FetchCounters::pad) and the
// host repeats the whole call in groups small enough for lists that hold the worst case (fetch_capi.cu).  Keeping the
// exact test out of this kernel -- round 1 decided overflowing leaves on the spot -- takes it from 168 to 76 registers;
// measured per 65,536 states at 384 / 512 / 768 threads per SM: 154 / 142 / 153 us.
constexpr int kEnvThreadsMax = 512, kEnvStack = 40;
#ifndef MPLB_ENV_CHUNK
#define MPLB_ENV_CHUNK 32
#endif
constexpr unsigned kEnvChunk = MPLB_ENV_CHUNK;
template <bool kSmemNodes>
__global__ void __launch_bounds__(kSmemNodes ? kEnvThreadsMax : 128)
fetchEnvKernel(FetchSceneDev sc, const double *__restrict__ frames, const unsigned *__restrict__ survivors,
               const unsigned *__restrict__ nSurvivors, const float2 *__restrict__ slack, unsigned char *__restrict__ touched,
               unsigned *__restrict__ envHit, WorkList leaves, unsigned *itemCounter, FetchCounters *ctr) {
    if (sc.nEnvNodes == 0) return;
    extern __shared__ __align__(16) unsigned char envSmem[];
    float4 *nodesS = reinterpret_cast<float4 *>(envSmem);
    if (kSmemNodes) {
        for (int i = threadIdx.x; i < sc.nEnvNodes * 4; i += blockDim.x) nodesS[(i >> 2) * 5 + (i & 3)] = __ldg(sc.envNodes + i);
        __syncthreads();
    }
    // per-geometry constants in shared memory: indexing the kernel parameters with a lane-varying geometry index
    // serialises in the constant cache (8 % of this kernel's samples)
    __shared__ float geomS[12][8];   // box half extents (x 1.00001), lift, reach
    if (threadIdx.x < 12) {
        float e[3];
        shapeExtents(sc.shapes[threadIdx.x], e);
        geomS[threadIdx.x][0] = e[0] * 1.00001f;
        geomS[threadIdx.x][1] = e[1] * 1.00001f;
        geomS[threadIdx.x][2] = e[2] * 1.00001f;
        geomS[threadIdx.x][3] = sc.lift[threadIdx.x];
        geomS[threadIdx.x][4] = sc.reach[threadIdx.x];
    }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned nItems = *nSurvivors * 12u;
    const float4 x0 = make_float4((float)sc.envFrame[0], (float)sc.envFrame[1], (float)sc.envFrame[2], (float)sc.envFrame[3]);
    const float4 x1 = make_float4((float)sc.envFrame[4], (float)sc.envFrame[5], (float)sc.envFrame[6], (float)sc.envFrame[7]);
    const float4 x2 = make_float4((float)sc.envFrame[8], (float)sc.envFrame[9], (float)sc.envFrame[10], (float)sc.envFrame[11]);
    // the per-lane stack lives in shared memory ([level][thread]: conflict free); as a local array it cost
    // 15 M local-memory sectors and 111 MB of DRAM write-back per 65,536-state launch (ncu)
    __shared__ int stackG[kSmemNodes ? 1 : kEnvStack][kSmemNodes ? 1 : 128];
    // [level][thread]; with the nodes in shared memory the stack follows them in the dynamic allocation
    int *stackBase = kSmemNodes ? reinterpret_cast<int *>(nodesS + (size_t)sc.nEnvNodes * 5) : &stackG[0][0];
    const unsigned tx = threadIdx.x, nThreads = blockDim.x;
    auto stackAt = [&](int level) -> int & { return stackBase[level * nThreads + tx]; };
    int sp = 0;
    // What depends on the shape's box and the (constant) environment frame only is formed once per item (device_math.cuh:
    // obbFrame); a stack entry >= 0 names a PAIR of sibling nodes (entry, entry + 1) whose parent is within reach -- both
    // are tested in one step, two independent chains of arithmetic per trip through the loop -- and ~node < 0 a single
    // node (an environment that is one triangle).  The root's own box is not tested: its children are, right away.
    ObbFrame F;
    unsigned long long state = 0;
    int g = 0;
    float reachSlack = kCullSlack;   // interval items: the box test is inflated by the geometry's displacement
    unsigned bv = 0;
    bool exhausted = false;
    unsigned chunkBase = 0, chunkLeft = 0;
    // The two counters every warp of the grid shares -- the item cursor and the leaf list's fill count -- are kept off the
    // loop's critical path (ncu: a quarter of this kernel's samples sat behind their atomics, waiting for the value to
    // come back from the one L2 slice that holds it):
    // * the next chunk is claimed while the current one still has items, its survivor indices are read one per lane and
    //   the frames they point to are prefetched, so that a lane that runs dry finds its item's frame in L1;
    // * leaves are collected in a per-warp shared-memory buffer and appended 32 or more at a time with one atomic.
    // Measured per 65,536 states: 136-148 -> 118-122 us.
    static_assert(kEnvChunk <= 32, "one item of a chunk per lane");
    unsigned chunkStart = 0, chunkState = 0;   // first item of the current chunk; survivors[] of item chunkStart + lane
    unsigned nextBase = 0;                     // (lane 0) the chunk claimed ahead
    bool nextClaimed = false;
    constexpr int kPushCap = 96;               // < 32 held between trips + at most 2 per lane and trip
    __shared__ unsigned long long pushBuf[(kSmemNodes ? kEnvThreadsMax : 128) / 32][kPushCap];
    __shared__ unsigned pushCnt[(kSmemNodes ? kEnvThreadsMax : 128) / 32];
    const unsigned wid = threadIdx.x >> 5;
    if (lane == 0) pushCnt[wid] = 0;
    __syncwarp();
    // The traversal culls against the LEAF'S BOX -- the triangle's bounding rectangle with some thickness.  Before a
    // buffered leaf is listed its triangle itself is tested against the shape's box, one leaf per lane (the whole warp
    // takes part: this is the one place where the kernel's lanes all have the same thing to do): a listed leaf costs the
    // exact kernel an MPR set-up and one or two supports -- a few thousand dependent FP64 instructions of a lane whose
    // warp waits for it -- and many of them are clear of the box all the same.
    auto flush = [&](unsigned held) {   // warp-collective
        for (unsigned i0 = 0; i0 < held; i0 += 32) {
            const unsigned i = i0 + lane;
            unsigned long long v = 0;
            bool keep = false;
            if (i < held) {
                v = pushBuf[wid][i];
                keep = true;
                if (MPLB_FETCH_TRI_SAT) {
                    const int gg = (int)((v >> 24) & 15);
                    FrameD fr;
                    loadFrame(frames + (v >> 28) * 144 + 12 * gg, fr);
                    const double *tp = sc.envTris64 + 9 * (size_t)(v & 0xffffff);
                    float q[3][3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float px = (float)tp[3 * k], py = (float)tp[3 * k + 1], pz = (float)tp[3 * k + 2];
                        // world coordinates relative to the box's centre, then onto its axes (the columns of R)
                        const float wx = fmaf(x0.x, px, fmaf(x0.y, py, fmaf(x0.z, pz, x0.w))) - (float)fr.t[0];
                        const float wy = fmaf(x1.x, px, fmaf(x1.y, py, fmaf(x1.z, pz, x1.w))) - (float)fr.t[1];
                        const float wz = fmaf(x2.x, px, fmaf(x2.y, py, fmaf(x2.z, pz, x2.w))) - (float)fr.t[2];
#pragma unroll
                        for (int a = 0; a < 3; ++a) q[k][a] = fmaf((float)fr.R[0][a], wx, fmaf((float)fr.R[1][a], wy, (float)fr.R[2][a] * wz));
                    }
                    const float e[3] = {geomS[gg][0], geomS[gg][1], geomS[gg][2]};
                    keep = !boxTriangleApart(e, q);
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (m) {
                unsigned at = 0;
                if (lane == 0) at = atomicAdd(leaves.count, (unsigned)__popc(m));
                at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
                if (keep) {
                    // list full: flagged, the host repeats the call with lists that hold the worst case
                    if (at < leaves.capacity) leaves.items[at] = v;
                    else atomicExch(&ctr->pad, 1u);
                }
            }
        }
        __syncwarp();
        if (lane == 0) pushCnt[wid] = 0;
        __syncwarp();
    };
    const int rootLink = __float_as_int(__ldg(sc.envNodes + 3).w);
    const int rootEntry = rootLink >= 0 ? rootLink : ~0;
    for (;;) {
        // refill the lanes whose traversal is over
        const unsigned need = __ballot_sync(0xffffffffu, sp == 0);
        if (need && !exhausted) {
            // items are claimed kEnvChunk at a time per warp and dealt out from registers: some lane runs dry in
            // nearly every iteration, and a global atomic per iteration sat on the loop's critical path
            if (chunkLeft == 0) {
                unsigned base = 0;
                if (lane == 0) base = nextClaimed ? nextBase : atomicAdd(itemCounter, kEnvChunk);
                nextClaimed = false;
                chunkBase = chunkStart = __shfl_sync(0xffffffffu, base, 0);
                chunkLeft = chunkBase < nItems ? min(kEnvChunk, nItems - chunkBase) : 0u;
                if (chunkLeft == 0) exhausted = true;
                if (lane < chunkLeft) {
                    const unsigned it = chunkBase + lane;
                    chunkState = survivors[it / 12u];
                    const double *fp = frames + (unsigned long long)chunkState * 144 + 12 * (it % 12u);
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(fp));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(fp + 11));   // (96 bytes: may straddle two lines)
                }
            }
            const unsigned rank = __popc(need & ((1u << lane) - 1u));
            const unsigned take = min((unsigned)__popc(need), chunkLeft);
            const unsigned base = chunkBase;
            chunkBase += take;
            chunkLeft -= take;
            const unsigned myState = __shfl_sync(0xffffffffu, chunkState, (base + rank - chunkStart) & 31u);
            if (chunkLeft < 12 && !nextClaimed && !exhausted) {   // claim ahead: the value is not needed before the chunk runs out
                if (lane == 0) nextBase = atomicAdd(itemCounter, kEnvChunk);
                nextClaimed = true;
            }
            if (sp == 0 && rank < take) {
                const unsigned item = base + rank;
                {
                    state = myState;
                    g = (int)(item % 12u);
                    FrameD fr;
                    loadFrame(frames + state * 144 + 12 * g, fr);
                    // shape box in world coordinates: axes = columns of R
                    const float4 a0 = make_float4((float)fr.R[0][0], (float)fr.R[1][0], (float)fr.R[2][0], (float)fr.t[0]);
                    const float4 a1 = make_float4((float)fr.R[0][1], (float)fr.R[1][1], (float)fr.R[2][1], (float)fr.t[1]);
                    const float4 a2 = make_float4((float)fr.R[0][2], (float)fr.R[1][2], (float)fr.R[2][2], (float)fr.t[2]);
                    obbFrame(a0, a1, a2, make_float4(geomS[g][0], geomS[g][1], geomS[g][2], 0), x0, x1, x2, F);
                    reachSlack = kCullSlack;
                    if (slack) {
                        const float2 sl = slack[state];
                        reachSlack += (sl.x * geomS[g][3] + sl.y * geomS[g][4]) * 1.0001f + 1e-6f;
                    }
                    stackAt(sp++) = rootEntry;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, sp > 0) == 0) {
            if (exhausted) break;
            continue;
        }
        if (sp > 0) {
            const int entry = stackAt(--sp);
            const bool pair = entry >= 0;
            const int node = pair ? entry : ~entry;
            float4 b0, b1, b2, be, c0, c1, c2, ce;
            if (kSmemNodes) {
                const float4 *pn = nodesS + 5 * (size_t)node;
                b0 = pn[0]; b1 = pn[1]; b2 = pn[2]; be = pn[3];
                const float4 *pm = pair ? pn + 5 : pn;
                c0 = pm[0]; c1 = pm[1]; c2 = pm[2]; ce = pm[3];
            } else {
                const float4 *pn = sc.envNodes + 4 * (size_t)node;
                ldg8(pn, b0, b1);
                ldg8(pn + 2, b2, be);
                const float4 *pm = pair ? pn + 4 : pn;
                ldg8(pm, c0, c1);
                ldg8(pm + 2, c2, ce);
            }
            bv += pair ? 2u : 1u;
            float dist2;
            const float gapB = obbGapFramed(F, b0, b1, b2, be, dist2);
            const float gapC = pair ? obbGapFramed(F, c0, c1, c2, ce, dist2) : 3.0e38f;
            // one survivor at a time: inner nodes hand their children pair to the stack, leaves go to the exact kernel
            auto survivor = [&](float gap, int link) {
                if (gap > reachSlack) return;
                if (link >= 0) {
                    if (sp + 1 > kEnvStack) {
                        atomicExch(&ctr->error, 1u);
                        sp = 0;
                    } else {
                        stackAt(sp++) = link;
                    }
                } else if (gap > kCullSlack) {
                    touched[state] |= 8;   // a triangle's box within reach of the interval, but not of this sample
                } else {
                    if (touched) touched[state] |= 16;
                    const unsigned tri = (unsigned)(-(link + 1));
                    pushBuf[wid][atomicAdd(&pushCnt[wid], 1u)] = (state << 28) | ((unsigned long long)g << 24) | tri;
                    if (MPLB_FETCH_TRI_SAT) {   // the flush reads the triangle (72 bytes: may straddle two lines)
                        const double *tp = sc.envTris64 + 9 * (size_t)tri;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(tp));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + 8));
                    }
                }
            };
            survivor(gapC, __float_as_int(ce.w));
            survivor(gapB, __float_as_int(be.w));
        }
        __syncwarp();
        const unsigned held = *(volatile unsigned *)&pushCnt[wid];
        if (held >= 32) flush(held);
    }
    {
        const unsigned held = *(volatile unsigned *)&pushCnt[wid];
        if (held) flush(held);
    }
    atomicAdd(&ctr->bvTests, (unsigned long long)bv);
}

// Kernel 4, one thread per listed (state, geometry, triangle): exact shape-triangle test
__global__ void __launch_bounds__(128)
fetchLeafKernel(FetchSceneDev sc, const double *__restrict__ frames, WorkList leaves, unsigned *envHit,
                FetchCounters *ctr) {
    const unsigned n = min(*leaves.count, leaves.capacity);
    unsigned tested = 0;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const unsigned long long v = leaves.items[j];
        const unsigned long long state = v >> 28;
        if (envHit[state]) continue;
        const int r = exactLeaf(sc, frames, state, (int)((v >> 24) & 15), (unsigned)(v & 0xffffff));
        ++tested;
        if (r < 0) atomicExch(&ctr->error, 1u);
        if (r == 1) envHit[state] = 1u;
    }
    if (tested) atomicAdd(&ctr->leafTests, (unsigned long long)tested);
}

__global__ void fetchCombineKernel(const unsigned char *__restrict__ selfHit, const unsigned *__restrict__ envHit,
                                   unsigned long long n, uint8_t *__restrict__ valid) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) valid[i] = (selfHit[i] || envHit[i]) ? 0 : 1;
}

// edge samples: any colliding sample kills its edge
__global__ void fetchEdgeReduceKernel(const unsigned char *__restrict__ selfHit, const unsigned *__restrict__ envHit,
                                      const unsigned long long *__restrict__ edgeItems, unsigned long long n,
                                      unsigned *__restrict__ dead, FetchCounters *ctr) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (selfHit[i] == 2) return;
    if (selfHit[i] || envHit[i]) dead[(unsigned)(edgeItems[i] >> 32)] = 1u;
}

// (Sorting the work lists by pair number / geometry before the exact kernels -- so that the lanes of a warp run the same
// shapes' support functions -- was built and measured: lanes per instruction 8.7 -> 10.0 (leaves) and 9.3 -> 12.2
// (pairs), 102 -> 80 us and 59 -> 51 us per 65,536 states, all of it spent again on the sort's own three small launches
// per list.  The divergence is in MPR's data-dependent iteration counts, not in the kind of shape.)
// the exact kernels' grids do not depend on the list lengths (grid-stride over the device-side counts)
cudaError_t runFetch(const FetchSceneDev &sc, const double *dStates, const double *dFrom, const double *dTo,
                     const unsigned *dSteps, const unsigned long long *dItems, const float2 *dSlack, const unsigned *dDead,
                     size_t n, const FetchScratch &w, FetchCounters *dCtr, int numSms, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(w.envHit, 0, n * sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    unsigned char *touched = dSlack ? w.touched : nullptr;
    if (touched) {
        e = cudaMemsetAsync(touched, 0, n, stream);
        if (e != cudaSuccess) return e;
    }
    e = cudaMemsetAsync(w.counts, 0, 12 * sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    WorkList pairs{w.pairItems, w.counts, (unsigned)std::min<size_t>(w.pairCap, 0xffffffffu)};
    WorkList leaves{w.leafItems, w.counts + 1, (unsigned)std::min<size_t>(w.leafCap, 0xffffffffu)};
    const int exactGrid = numSms * 8;
    fetchFkKernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(sc, dStates, dFrom, dTo, dSteps, dItems, dSlack, dDead, n,
                                                                   w.frames, w.selfHit, touched, pairs, dCtr);
    // counts: [4] pairs that need MPR, [5] / [6] item cursors of the two fetchMprKernel launches
    WorkList exactPairs{w.pairItems + pairs.capacity, w.counts + 4, pairs.capacity};
    // (measured per 65,536 states: pairs decided inside fetchPairKernel 55-59 us; listed and run through fetchMprKernel
    // 15 + 35 us since that kernel prefetches its items (18 + 52-58 us before: a pair item is two object set-ups for one
    // or two supports); the shape-triangle tests: 102 -> 80 us.  MPLB_FETCH_INLINE_PAIRS=1 / MPLB_FETCH_INLINE_MPR=1
    // select the other ways.)
    static const bool inlineMpr = std::getenv("MPLB_FETCH_INLINE_MPR") != nullptr;
    static const bool mprPairs = std::getenv("MPLB_FETCH_INLINE_PAIRS") == nullptr && !inlineMpr;
    fetchPairKernel<<<exactGrid, 128, 0, stream>>>(sc, w.frames, pairs, mprPairs ? exactPairs : WorkList{nullptr, w.counts + 4, 0u},
                                                   dSlack, touched, w.selfHit, dCtr);
    if (mprPairs) fetchMprKernel<false><<<exactGrid, 128, 0, stream>>>(sc, w.frames, exactPairs, w.counts + 5, w.selfHit, w.envHit, dCtr);
    fetchSurvivorsKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w.selfHit, n, w.survivors, w.counts + 2);
    // Shared-memory variant: as many warps (<= 512 threads) as the hierarchy leaves room for next to the per-thread
    // stacks.  Small groups (edge rounds, scalar calls) take the global-memory variant: staging the hierarchy into every
    // SM's shared memory -- and the carve-out switch the 160-odd KB ask for -- costs ~45 us before the first box test.
    static const size_t smemFrom = [] {
        const char *e = std::getenv("MPLB_ENV_SMEM_FROM");   // diagnostic: group size from which the hierarchy is staged
        return e ? (size_t)std::atol(e) : (size_t)4096;
    }();
    int envThreads = 0;
    for (int t : {512, 384, 256})
        if (!envThreads && n >= smemFrom && (size_t)sc.nEnvNodes * 80 + (size_t)kEnvStack * t * sizeof(int) <= 210 * 1024) envThreads = t;   // (+ 13 KB of static shared memory)
    if (envThreads) {
        const size_t envSmemBytes = (size_t)sc.nEnvNodes * 80 + (size_t)kEnvStack * envThreads * sizeof(int);
        e = cudaFuncSetAttribute(fetchEnvKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)envSmemBytes);
        if (e != cudaSuccess) return e;
        fetchEnvKernel<true><<<numSms, envThreads, envSmemBytes, stream>>>(sc, w.frames, w.survivors, w.counts + 2, dSlack, touched,
                                                                          w.envHit, leaves, w.counts + 3, dCtr);
    } else {
        const int grid = (int)std::min<size_t>((size_t)exactGrid, (n * 12 + 127) / 128);
        fetchEnvKernel<false><<<std::max(grid, 1), 128, 0, stream>>>(sc, w.frames, w.survivors, w.counts + 2, dSlack, touched, w.envHit,
                                                                     leaves, w.counts + 3, dCtr);
    }
    if (inlineMpr) fetchLeafKernel<<<exactGrid, 128, 0, stream>>>(sc, w.frames, leaves, w.envHit, dCtr);
    else fetchMprKernel<true><<<exactGrid, 128, 0, stream>>>(sc, w.frames, leaves, w.counts + 6, w.selfHit, w.envHit, dCtr);
    return cudaGetLastError();
}

// ---- interval bookkeeping of edge batches on the device --------------------------------------------------------------
// fetch_capi.cu used to keep the interval lists on the host: per round an OpenMP pass to build the items, two copies
// down, two copies up and another pass to split what could not be certified -- half of an edge batch's time.  Here the
// lists stay in device memory; per round the host launches the groups and reads one count back.
// An interval is (edge, lo, hi, -): samples lo..hi of the edge, lo == hi == steps standing for `to` itself.

// how far the torso (prismatic) and the arm joints (L1) move over the whole edge, rounded up
__global__ void fetchEdgeMotionKernel(const double *__restrict__ from, const double *__restrict__ to, unsigned n,
                                      float2 *__restrict__ motion) {
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double a = 0;
    for (int j = 1; j < 8; ++j) a += fabs(to[8 * (size_t)e + j] - from[8 * (size_t)e + j]);
    motion[e] = make_float2((float)(fabs(to[8 * (size_t)e] - from[8 * (size_t)e]) * 1.000001), (float)(a * 1.000001));
}
// round 0: `to` and up to 16 equal intervals of the samples [1, steps - 1] of every edge
__global__ void fetchRound0Kernel(const unsigned *__restrict__ steps, unsigned n, uint4 *__restrict__ list, unsigned *count,
                                  unsigned cap, unsigned *overflow) {
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned st = steps[e];
    const unsigned long long m = st ? st - 1 : 0;
    unsigned k = 1;
    for (unsigned long long j = 0; j < 16; ++j) k += (1u + (unsigned)((j * m) >> 4)) <= (unsigned)(((j + 1) * m) >> 4) ? 1u : 0u;
    unsigned at = atomicAdd(count, k);
    if (at + k > cap) {
        atomicExch(overflow, 1u);
        return;
    }
    list[at++] = make_uint4(e, st, st, 0);
    for (unsigned long long j = 0; j < 16; ++j) {
        const unsigned lo = 1u + (unsigned)((j * m) >> 4), hi = (unsigned)(((j + 1) * m) >> 4);
        if (lo <= hi) list[at++] = make_uint4(e, lo, hi, 0);
    }
}
// one group of intervals -> the items / slacks runFetch takes.  An interval over which the gripper can stray by more
// than `cap` is practically never certified and its inflated traversal costs more than a plain test: its middle sample
// is evaluated plainly (slack 0) and it is split unconditionally (forced).
__global__ void fetchItemsKernel(const uint4 *__restrict__ list, unsigned cnt, const unsigned *__restrict__ steps,
                                 const float2 *__restrict__ motion, float cap, unsigned long long *__restrict__ items,
                                 float2 *__restrict__ slack, unsigned char *__restrict__ forced) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const uint4 v = list[i];
    const unsigned st = steps[v.x], mid = v.y + (v.z - v.y) / 2;
    items[i] = ((unsigned long long)v.x << 32) | mid;
    float f = 0.f;
    if (v.z > v.y) f = (float)((double)max(v.z - mid, mid - v.y) / (double)st * 1.000001);
    const float2 mo = motion[v.x];
    float2 sl = make_float2(mo.x * f, mo.y * f);
    const bool fo = v.z > v.y && sl.x + sl.y * 1.2f > cap;
    forced[i] = fo ? 1 : 0;
    slack[i] = fo ? make_float2(0.f, 0.f) : sl;
}
// after the group's tests: what is left of every interval that was neither certified nor killed goes to the next round
__global__ void fetchSplitKernel(const uint4 *__restrict__ list, unsigned cnt, const unsigned char *__restrict__ touched,
                                 const unsigned char *__restrict__ forced, const unsigned *__restrict__ dead,
                                 const unsigned *__restrict__ steps, const float2 *__restrict__ motion, float cap,
                                 unsigned minParts, uint4 *__restrict__ next, unsigned *nextCount, unsigned nextCap,
                                 unsigned *overflow) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const uint4 v = list[i];
    if (dead[v.x] || v.z <= v.y || !(touched[i] || forced[i])) return;   // dead, a single sample, or certified
    const unsigned mid = v.y + (v.z - v.y) / 2;
    // a split that was forced goes straight to pieces short enough to stand a chance (<= 2 H + 1 samples, H the half width
    // at which the gripper's stray reaches the cap), not through log2 plain halvings
    unsigned long long piece = 0xffffffffull;
    if (forced[i]) {
        const float2 mo = motion[v.x];
        const double perSample = ((double)mo.y * 1.2 + (double)mo.x) / (double)steps[v.x];
        const double H = perSample > 0 ? floor((double)cap / perSample) : 1e9;
        piece = (unsigned long long)fmin(1e9, 2.0 * fmax(H, 1.0) + 1.0);
    }
    unsigned long long parts[2] = {0, 0}, len[2] = {0, 0};
    unsigned lo[2] = {v.y, mid + 1};
    // (minParts > 1: a round with few intervals left is all latency -- eight kernels and a read-back for a few thousand
    // items -- so each half is cut into several pieces at once and the rounds that remain are fewer)
    if (mid > v.y) {
        len[0] = (unsigned long long)(mid - 1) - v.y + 1;
        parts[0] = max((len[0] + piece - 1) / piece, min((unsigned long long)minParts, len[0]));
    }
    if (mid < v.z) {
        len[1] = (unsigned long long)v.z - (mid + 1) + 1;
        parts[1] = max((len[1] + piece - 1) / piece, min((unsigned long long)minParts, len[1]));
    }
    const unsigned long long total = parts[0] + parts[1];
    if (total == 0) return;
    if (total > 0x7fffffffull) {
        atomicExch(overflow, 1u);
        return;
    }
    unsigned at = atomicAdd(nextCount, (unsigned)total);
    if ((unsigned long long)at + total > nextCap) {
        atomicExch(overflow, 1u);
        return;
    }
    for (int h = 0; h < 2; ++h)
        for (unsigned long long p = 0; p < parts[h]; ++p)
            next[at++] = make_uint4(v.x, lo[h] + (unsigned)(p * len[h] / parts[h]), lo[h] + (unsigned)((p + 1) * len[h] / parts[h]) - 1, 0);
}

}  // namespace

cudaError_t launchFetchEdgeMotion(const double *dFrom, const double *dTo, size_t n, float2 *dMotion, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    fetchEdgeMotionKernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(dFrom, dTo, (unsigned)n, dMotion);
    return cudaGetLastError();
}
cudaError_t launchFetchRound0(const unsigned *dSteps, size_t n, uint4 *dList, unsigned *dCount, unsigned cap, unsigned *dOverflow,
                              cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    fetchRound0Kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(dSteps, (unsigned)n, dList, dCount, cap, dOverflow);
    return cudaGetLastError();
}
cudaError_t launchFetchItems(const uint4 *dList, size_t cnt, const unsigned *dSteps, const float2 *dMotion, float cap,
                             unsigned long long *dItems, float2 *dSlack, unsigned char *dForced, cudaStream_t stream) {
    if (cnt == 0) return cudaSuccess;
    fetchItemsKernel<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(dList, (unsigned)cnt, dSteps, dMotion, cap, dItems, dSlack, dForced);
    return cudaGetLastError();
}
cudaError_t launchFetchSplit(const uint4 *dList, size_t cnt, const unsigned char *dTouched, const unsigned char *dForced,
                             const unsigned *dDead, const unsigned *dSteps, const float2 *dMotion, float cap, unsigned minParts,
                             uint4 *dNext, unsigned *dNextCount, unsigned nextCap, unsigned *dOverflow, cudaStream_t stream) {
    if (cnt == 0) return cudaSuccess;
    fetchSplitKernel<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(dList, (unsigned)cnt, dTouched, dForced, dDead, dSteps, dMotion,
                                                                        cap, minParts, dNext, dNextCount, nextCap, dOverflow);
    return cudaGetLastError();
}

cudaError_t launchFetchStates(const FetchSceneDev &sc, const double *dStates, size_t n, const FetchScratch &w,
                              uint8_t *dValid, FetchCounters *dCtr, int numSms, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (n >= (1ull << 36)) return cudaErrorInvalidValue;   // work descriptors keep 36 bits for the state index
    cudaError_t e = runFetch(sc, dStates, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, n, w, dCtr, numSms, stream);
    if (e != cudaSuccess) return e;
    fetchCombineKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w.selfHit, w.envHit, n, dValid);
    return cudaGetLastError();
}

cudaError_t launchFetchEdgeItems(const FetchSceneDev &sc, const double *dFrom, const double *dTo, const unsigned *dSteps,
                                 const unsigned long long *dItems, const float2 *dSlack, size_t nItems, unsigned *dDead,
                                 const FetchScratch &w, FetchCounters *dCtr, int numSms, cudaStream_t stream) {
    if (nItems == 0) return cudaSuccess;
    cudaError_t e = runFetch(sc, nullptr, dFrom, dTo, dSteps, dItems, dSlack, dDead, nItems, w, dCtr, numSms, stream);
    if (e != cudaSuccess) return e;
    fetchEdgeReduceKernel<<<(unsigned)((nItems + 255) / 256), 256, 0, stream>>>(w.selfHit, w.envHit, dItems, nItems, dDead, dCtr);
    return cudaGetLastError();
}

}  // namespace mplb
