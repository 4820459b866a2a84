// Device helpers shared by the validity kernels: FP64 "specification" arithmetic without FMA
// contraction (identical operation order to oracle/mplb_oracle.c) and the conservative FP32
// OBB-OBB separating-axis test.
#pragma once
#include <cuda_runtime.h>

namespace mplb {

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }
// two consecutive float4 (32-byte aligned) with one 256-bit load (sm_100: LDG.E.256): one L1 tag lookup
// per 32-byte sector instead of two -- the box-pair step is bound by L1TEX lookups, not by bytes
__device__ __forceinline__ void ldg8(const float4 *p, float4 &a, float4 &b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

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

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float dot(F3 a, F3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ F3 cross(F3 a, F3 b) {
    return {fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x)};
}

// OBB-OBB separating-axis test (Gottschalk's 15 axes) as a gap: the largest amount by which any of the
// axes separates the boxes (negative: every axis overlaps).  A and B are boxes in their own frames,
// (R,T) = rows x0..x2 maps B's frame into A's.  The boxes overlap for a caller iff gap <= slack, with a
// slack above the FP32 evaluation error (so a pair is only culled when the exact boxes are disjoint)
// plus, for interval certificates, the reach of the motion.  Cross-product axes are not normalised
// (|axis| <= 1), which only makes their gap smaller, i.e. the test more conservative.
__device__ __forceinline__ float obbGap(const float4 a0, const float4 a1, const float4 a2, const float4 ae,
                                        const float4 b0, const float4 b1, const float4 b2, const float4 be,
                                        const float4 x0, const float4 x1, const float4 x2, float &dist2) {
    // B's centre in A's frame, relative to A's centre, in A's box axes
    float cx = fmaf(x0.x, b0.w, fmaf(x0.y, b1.w, fmaf(x0.z, b2.w, x0.w))) - a0.w;
    float cy = fmaf(x1.x, b0.w, fmaf(x1.y, b1.w, fmaf(x1.z, b2.w, x1.w))) - a1.w;
    float cz = fmaf(x2.x, b0.w, fmaf(x2.y, b1.w, fmaf(x2.z, b2.w, x2.w))) - a2.w;
    dist2 = fmaf(cx, cx, fmaf(cy, cy, cz * cz));
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
    float gap = -3.0e38f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        gap = fmaxf(gap, fabsf(T[i]) - (ea[i] + fmaf(eb[0], Bf[i][0], fmaf(eb[1], Bf[i][1], eb[2] * Bf[i][2]))));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float s = fmaf(T[0], B[0][j], fmaf(T[1], B[1][j], T[2] * B[2][j]));
        gap = fmaxf(gap, fabsf(s) - (eb[j] + fmaf(ea[0], Bf[0][j], fmaf(ea[1], Bf[1][j], ea[2] * Bf[2][j]))));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            float s = fmaf(T[i2], B[i1][j], -T[i1] * B[i2][j]);
            float r = fmaf(ea[i1], Bf[i2][j], fmaf(ea[i2], Bf[i1][j], fmaf(eb[j1], Bf[i][j2], eb[j2] * Bf[i][j1])));
            gap = fmaxf(gap, fabsf(s) - r);
        }
    }
    return gap;
}


// The same test split in two, for a step that tests both children of one box against the same fixed box A: what
// depends on A and the transform only -- M = A (R), t0 = A (T - centre_A), both in A's box axes -- is formed once
// (39 operations), each child then costs B = M b^T (27), its centre M c_b + t0 (9) and the 15 axes.  Same quantities
// as obbGap up to rounding (a few eps of the coordinates' magnitude, well inside the callers' slack).
struct ObbFrame {
    float M[3][3], t0[3], ea[3];
};
__device__ __forceinline__ void obbFrame(const float4 a0, const float4 a1, const float4 a2, const float4 ae,
                                         const float4 x0, const float4 x1, const float4 x2, ObbFrame &F) {
    const float4 a[3] = {a0, a1, a2};
    const float dx = x0.w - a0.w, dy = x1.w - a1.w, dz = x2.w - a2.w;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        F.M[i][0] = fmaf(a[i].x, x0.x, fmaf(a[i].y, x1.x, a[i].z * x2.x));
        F.M[i][1] = fmaf(a[i].x, x0.y, fmaf(a[i].y, x1.y, a[i].z * x2.y));
        F.M[i][2] = fmaf(a[i].x, x0.z, fmaf(a[i].y, x1.z, a[i].z * x2.z));
        F.t0[i] = fmaf(a[i].x, dx, fmaf(a[i].y, dy, a[i].z * dz));
    }
    F.ea[0] = ae.x; F.ea[1] = ae.y; F.ea[2] = ae.z;
}
__device__ __forceinline__ float obbGapFramed(const ObbFrame &F, const float4 b0, const float4 b1, const float4 b2,
                                              const float4 be, float &dist2) {
    float T[3], B[3][3], Bf[3][3];
    const float4 b[3] = {b0, b1, b2};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        T[i] = fmaf(F.M[i][0], b0.w, fmaf(F.M[i][1], b1.w, fmaf(F.M[i][2], b2.w, F.t0[i])));
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            B[i][j] = fmaf(F.M[i][0], b[j].x, fmaf(F.M[i][1], b[j].y, F.M[i][2] * b[j].z));
            Bf[i][j] = fabsf(B[i][j]);
        }
    }
    dist2 = fmaf(T[0], T[0], fmaf(T[1], T[1], T[2] * T[2]));
    const float *ea = F.ea;
    const float eb[3] = {be.x, be.y, be.z};
    float gap = -3.0e38f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        gap = fmaxf(gap, fabsf(T[i]) - (ea[i] + fmaf(eb[0], Bf[i][0], fmaf(eb[1], Bf[i][1], eb[2] * Bf[i][2]))));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float s = fmaf(T[0], B[0][j], fmaf(T[1], B[1][j], T[2] * B[2][j]));
        gap = fmaxf(gap, fabsf(s) - (eb[j] + fmaf(ea[0], Bf[0][j], fmaf(ea[1], Bf[1][j], ea[2] * Bf[2][j]))));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            float s = fmaf(T[i2], B[i1][j], -T[i1] * B[i2][j]);
            float r = fmaf(ea[i1], Bf[i2][j], fmaf(ea[i2], Bf[i1][j], fmaf(eb[j1], Bf[i][j2], eb[j2] * Bf[i][j1])));
            gap = fmaxf(gap, fabsf(s) - r);
        }
    }
    return gap;
}

__device__ __forceinline__ bool obbOverlap(const float4 a0, const float4 a1, const float4 a2, const float4 ae,
                                           const float4 b0, const float4 b1, const float4 b2, const float4 be,
                                           const float4 x0, const float4 x1, const float4 x2, float slack, float &dist2) {
    return obbGap(a0, a1, a2, ae, b0, b1, b2, be, x0, x1, x2, dist2) <= slack;
}

}  // namespace mplb
