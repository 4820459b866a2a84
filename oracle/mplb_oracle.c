/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see mplb_oracle.h).  PARITY UNPINNED (see mplb_oracle.h).
 *
 * FP64 restatement of the SE(3) rigid-body validity path:
 *   se3_rigid_body_scenario.hpp:44-47   stateToTransform  (Eigen Quaternion::toRotationMatrix [EXT])
 *   se3_rigid_body_scenario.hpp:119-126 isValid(q)  -> fcl::collide(robot, tf, env, I) [EXT FCL]
 *   se3_rigid_body_scenario.hpp:134-178 isValid(from,to)  (bisection ring queue, 1024 slots)
 *   interpolate.hpp:9-36                lerp / Eigen slerp [EXT]
 *   nigh SE3Space / L1Space metrics [EXT]
 * FCL pieces restated ([EXT]): BVHModel<OBBRSS> top-down build with
 * SPLIT_METHOD_MEAN, obbDisjoint (15-axis, +1e-6), collisionRecurse with firstOverSecond and
 * canStop after the first contact, Intersect::intersect_Triangle (17 x project6).
 *
 * Build: gcc -O3 -march=native -ffp-contract=off -fopenmp (oracle/Makefile).
 */
#include "mplb_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double x, y, z; } v3;

static inline v3 v3_make(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline double v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 v3_cross(v3 a, v3 b) {
    return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

/* ------------------------------------------------------------------------------------------ */
/* OBB BVH, FCL-like ([EXT] BVHModel::recursiveBuildTree, BVFitter<OBB>, BVSplitter mean)      */

typedef struct {
    double axis[3][3]; /* axis[k] = k-th box axis (unit, world/model frame) */
    double center[3];
    double extent[3];
    int first_child;   /* >=0: children first_child, first_child+1; <0: leaf, tri = -(fc+1) */
} obb_node;

struct orc_mesh {
    size_t ntris;
    v3 *tris; /* 3*ntris vertices */
    size_t nnodes;
    obb_node *nodes;
};

static void jacobi_eigen3(double A[3][3], double evals[3], double evecs[3][3]) {
    /* cyclic Jacobi for a symmetric 3x3; evecs columns are eigenvectors */
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double a[3][3];
    memcpy(a, A, sizeof(a));
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(a[p][q]) < 1e-300) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) evals[i] = a[i][i];
    memcpy(evecs, V, sizeof(V));
}

static void fit_obb(const v3 *tris, const int *prim, int n, obb_node *nd) {
    /* covariance of the 3n triangle vertices ([EXT] getCovariance) */
    double S1[3] = {0, 0, 0}, S2[3][3] = {{0}};
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            v3 p = tris[3 * prim[i] + k];
            double v[3] = {p.x, p.y, p.z};
            for (int a = 0; a < 3; ++a) {
                S1[a] += v[a];
                for (int b = 0; b < 3; ++b) S2[a][b] += v[a] * v[b];
            }
        }
    double np = 3.0 * n, M[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) M[a][b] = S2[a][b] - S1[a] * S1[b] / np;
    double ev[3], E[3][3];
    jacobi_eigen3(M, ev, E);
    /* [EXT] axisFromEigen: axis0 = largest, axis1 = middle, axis2 = axis0 x axis1 */
    int mx = 0, mn = 0;
    for (int i = 1; i < 3; ++i) {
        if (ev[i] > ev[mx]) mx = i;
        if (ev[i] < ev[mn]) mn = i;
    }
    int mid = 3 - mx - mn;
    if (mx == mn) { mx = 0; mid = 1; }
    v3 a0 = v3_make(E[0][mx], E[1][mx], E[2][mx]);
    v3 a1 = v3_make(E[0][mid], E[1][mid], E[2][mid]);
    double l0 = sqrt(v3_dot(a0, a0)), l1 = sqrt(v3_dot(a1, a1));
    a0 = v3_make(a0.x / l0, a0.y / l0, a0.z / l0);
    a1 = v3_make(a1.x / l1, a1.y / l1, a1.z / l1);
    v3 a2 = v3_cross(a0, a1);
    v3 ax[3] = {a0, a1, a2};
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            v3 p = tris[3 * prim[i] + k];
            for (int a = 0; a < 3; ++a) {
                double d = v3_dot(ax[a], p);
                if (d < lo[a]) lo[a] = d;
                if (d > hi[a]) hi[a] = d;
            }
        }
    double c[3], e[3];
    for (int a = 0; a < 3; ++a) {
        c[a] = 0.5 * (lo[a] + hi[a]);
        e[a] = 0.5 * (hi[a] - lo[a]);
    }
    for (int a = 0; a < 3; ++a) {
        nd->axis[a][0] = ax[a].x; nd->axis[a][1] = ax[a].y; nd->axis[a][2] = ax[a].z;
        nd->extent[a] = e[a];
    }
    for (int d = 0; d < 3; ++d)
        nd->center[d] = c[0] * nd->axis[0][d] + c[1] * nd->axis[1][d] + c[2] * nd->axis[2][d];
}

static void build_rec(orc_mesh *m, int *prim, int node, int first, int n) {
    obb_node *nd = &m->nodes[node];
    fit_obb(m->tris, prim + first, n, nd);
    if (n == 1) {
        nd->first_child = -(prim[first] + 1);
        return;
    }
    /* SPLIT_METHOD_MEAN: axis 0, value = mean projected centroid */
    v3 ax = v3_make(nd->axis[0][0], nd->axis[0][1], nd->axis[0][2]);
    double sum = 0;
    for (int i = 0; i < n; ++i) {
        const v3 *t = &m->tris[3 * prim[first + i]];
        sum += v3_dot(ax, v3_add(v3_add(t[0], t[1]), t[2]));
    }
    double split = sum / (3.0 * n);
    int c1 = 0;
    for (int i = 0; i < n; ++i) {
        const v3 *t = &m->tris[3 * prim[first + i]];
        v3 c = v3_add(v3_add(t[0], t[1]), t[2]);
        double proj = v3_dot(ax, v3_make(c.x / 3.0, c.y / 3.0, c.z / 3.0));
        if (!(proj > split)) {
            int tmp = prim[first + i]; prim[first + i] = prim[first + c1]; prim[first + c1] = tmp;
            ++c1;
        }
    }
    if (c1 == 0 || c1 == n) c1 = n / 2;
    int fc = (int)m->nnodes;
    m->nnodes += 2;
    nd->first_child = fc;
    build_rec(m, prim, fc, first, c1);
    build_rec(m, prim, fc + 1, first + c1, n - c1);
}

orc_mesh *orc_mesh_create(const double *tris, size_t n) {
    orc_mesh *m = (orc_mesh *)calloc(1, sizeof(orc_mesh));
    m->ntris = n;
    m->tris = (v3 *)malloc(sizeof(v3) * 3 * (n ? n : 1));
    memcpy(m->tris, tris, sizeof(double) * 9 * n);
    if (n == 0) return m;
    m->nodes = (obb_node *)malloc(sizeof(obb_node) * (2 * n - 1));
    int *prim = (int *)malloc(sizeof(int) * n);
    for (size_t i = 0; i < n; ++i) prim[i] = (int)i;
    m->nnodes = 1;
    build_rec(m, prim, 0, 0, (int)n);
    free(prim);
    return m;
}
void orc_mesh_destroy(orc_mesh *m) {
    if (!m) return;
    free(m->tris); free(m->nodes); free(m);
}
size_t orc_mesh_num_tris(const orc_mesh *m) { return m->ntris; }
size_t orc_mesh_num_nodes(const orc_mesh *m) { return m->nnodes; }

/* ------------------------------------------------------------------------------------------ */
/* pose -> relative transform                                                                  */

typedef struct { double R[3][3]; double T[3]; } xform;

/* Eigen QuaternionBase::toRotationMatrix [EXT], coefficient order x,y,z,w, no normalisation */
static void quat_to_R(const double q[4], double R[3][3]) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w;
    double txx = tx * x, txy = ty * x, txz = tz * x;
    double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz;       R[0][2] = txz + twy;
    R[1][0] = txy + twz;       R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
    R[2][0] = txz - twy;       R[2][1] = tyz + twx;       R[2][2] = 1 - (txx + tyy);
}

/* fcl::collide(robot, tf1, env, Identity): env expressed in the robot frame,
 * R = R1^T, T = R1^T (0 - t1)  ([EXT] relativeTransform). */
static void state_to_rel(const double s[7], xform *X) {
    double R1[3][3];
    quat_to_R(s, R1);
    double nt[3] = {0.0 - s[4], 0.0 - s[5], 0.0 - s[6]};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) X->R[i][j] = R1[j][i];
        X->T[i] = (R1[0][i] * nt[0] + R1[1][i] * nt[1]) + R1[2][i] * nt[2];
    }
}

static inline v3 xf_apply(const xform *X, v3 p) {
    return v3_make(((X->R[0][0] * p.x + X->R[0][1] * p.y) + X->R[0][2] * p.z) + X->T[0],
                   ((X->R[1][0] * p.x + X->R[1][1] * p.y) + X->R[1][2] * p.z) + X->T[1],
                   ((X->R[2][0] * p.x + X->R[2][1] * p.y) + X->R[2][2] * p.z) + X->T[2]);
}

/* ------------------------------------------------------------------------------------------ */
/* triangle-triangle: [EXT] fcl::detail::Intersect::intersect_Triangle / project6             */

static inline int project6(v3 ax, v3 p1, v3 p2, v3 p3, v3 q1, v3 q2, v3 q3) {
    double P1 = v3_dot(ax, p1), P2 = v3_dot(ax, p2), P3 = v3_dot(ax, p3);
    double Q1 = v3_dot(ax, q1), Q2 = v3_dot(ax, q2), Q3 = v3_dot(ax, q3);
    double mx1 = fmax(P1, fmax(P2, P3)), mn1 = fmin(P1, fmin(P2, P3));
    double mx2 = fmax(Q1, fmax(Q2, Q3)), mn2 = fmin(Q1, fmin(Q2, Q3));
    if (mn1 > mx2) return 0;
    if (mn2 > mx1) return 0;
    return 1;
}

/* P* in the robot frame, Q* already transformed into the robot frame */
static int tri_tri_intersect(v3 P1, v3 P2, v3 P3, v3 Q1, v3 Q2, v3 Q3) {
    v3 p1 = v3_sub(P1, P1), p2 = v3_sub(P2, P1), p3 = v3_sub(P3, P1);
    v3 q1 = v3_sub(Q1, P1), q2 = v3_sub(Q2, P1), q3 = v3_sub(Q3, P1);
    v3 e1 = v3_sub(p2, p1), e2 = v3_sub(p3, p2), e3 = v3_sub(p1, p3);
    v3 f1 = v3_sub(q2, q1), f2 = v3_sub(q3, q2), f3 = v3_sub(q1, q3);
    v3 n1 = v3_cross(e1, e2), m1 = v3_cross(f1, f2);
    if (!project6(n1, p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(m1, p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e1, f1), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e1, f2), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e1, f3), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e2, f1), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e2, f2), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e2, f3), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e3, f1), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e3, f2), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e3, f3), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e1, n1), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e2, n1), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(e3, n1), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(f1, m1), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(f2, m1), p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(v3_cross(f3, m1), p1, p2, p3, q1, q2, q3)) return 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* OBB-OBB: [EXT] fcl::overlap(R0,T0,b1,b2) + obbDisjoint (Gottschalk, reps = 1e-6)           */

static int obb_disjoint(const double B[3][3], const double T[3], const double a[3],
                        const double b[3]) {
    const double reps = 1e-6;
    double Bf[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Bf[i][j] = fabs(B[i][j]) + reps;
    double t, s;
    for (int i = 0; i < 3; ++i) { /* A axes */
        t = fabs(T[i]);
        if (t > a[i] + ((b[0] * Bf[i][0] + b[1] * Bf[i][1]) + b[2] * Bf[i][2])) return 1;
    }
    for (int j = 0; j < 3; ++j) { /* B axes */
        s = (T[0] * B[0][j] + T[1] * B[1][j]) + T[2] * B[2][j];
        t = fabs(s);
        if (t > b[j] + ((a[0] * Bf[0][j] + a[1] * Bf[1][j]) + a[2] * Bf[2][j])) return 1;
    }
    for (int i = 0; i < 3; ++i) { /* A_i x B_j */
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        for (int j = 0; j < 3; ++j) {
            int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            s = T[i2] * B[i1][j] - T[i1] * B[i2][j];
            t = fabs(s);
            if (t > (a[i1] * Bf[i2][j] + a[i2] * Bf[i1][j]) + (b[j1] * Bf[i][j2] + b[j2] * Bf[i][j1]))
                return 1;
        }
    }
    return 0;
}

static int obb_overlap(const xform *X, const obb_node *b1, const obb_node *b2) {
    /* Ttemp = R0*b2.To + T0 - b1.To ; T = b1.axis^T Ttemp ; R = b1.axis^T R0 b2.axis */
    v3 c2 = xf_apply(X, v3_make(b2->center[0], b2->center[1], b2->center[2]));
    double Tt[3] = {c2.x - b1->center[0], c2.y - b1->center[1], c2.z - b1->center[2]};
    double T[3], RA2[3][3], R[3][3];
    for (int i = 0; i < 3; ++i)
        T[i] = (b1->axis[i][0] * Tt[0] + b1->axis[i][1] * Tt[1]) + b1->axis[i][2] * Tt[2];
    for (int i = 0; i < 3; ++i)     /* RA2[:,j] = R0 * axis2_j */
        for (int j = 0; j < 3; ++j)
            RA2[i][j] = (X->R[i][0] * b2->axis[j][0] + X->R[i][1] * b2->axis[j][1]) +
                        X->R[i][2] * b2->axis[j][2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            R[i][j] = (b1->axis[i][0] * RA2[0][j] + b1->axis[i][1] * RA2[1][j]) +
                      b1->axis[i][2] * RA2[2][j];
    return !obb_disjoint(R, T, b1->extent, b2->extent);
}

/* ------------------------------------------------------------------------------------------ */

int orc_se3_collide_brute(const orc_mesh *robot, const orc_mesh *env, const double state[7]) {
    xform X;
    state_to_rel(state, &X);
    for (size_t j = 0; j < env->ntris; ++j) {
        v3 Q1 = xf_apply(&X, env->tris[3 * j]), Q2 = xf_apply(&X, env->tris[3 * j + 1]),
           Q3 = xf_apply(&X, env->tris[3 * j + 2]);
        for (size_t i = 0; i < robot->ntris; ++i) {
            const v3 *P = &robot->tris[3 * i];
            if (tri_tri_intersect(P[0], P[1], P[2], Q1, Q2, Q3)) return 1;
        }
    }
    return 0;
}

typedef struct {
    const orc_mesh *m1, *m2;
    xform X;
    int hit;
    uint64_t bv, tri;
} trav;

static inline double obb_size(const obb_node *n) {
    return (n->extent[0] * n->extent[0] + n->extent[1] * n->extent[1]) + n->extent[2] * n->extent[2];
}

/* [EXT] fcl::detail::collisionRecurse */
static void collision_recurse(trav *t, int b1, int b2) {
    const obb_node *n1 = &t->m1->nodes[b1], *n2 = &t->m2->nodes[b2];
    int l1 = n1->first_child < 0, l2 = n2->first_child < 0;
    t->bv++;
    if (!obb_overlap(&t->X, n1, n2)) return;
    if (l1 && l2) {
        const v3 *P = &t->m1->tris[3 * (-(n1->first_child) - 1)];
        const v3 *Q = &t->m2->tris[3 * (-(n2->first_child) - 1)];
        t->tri++;
        if (tri_tri_intersect(P[0], P[1], P[2], xf_apply(&t->X, Q[0]), xf_apply(&t->X, Q[1]),
                              xf_apply(&t->X, Q[2])))
            t->hit = 1;
        return;
    }
    if (l2 || (!l1 && obb_size(n1) > obb_size(n2))) {
        collision_recurse(t, n1->first_child, b2);
        if (t->hit) return; /* canStop(): num_max_contacts = 1 */
        collision_recurse(t, n1->first_child + 1, b2);
    } else {
        collision_recurse(t, b1, n2->first_child);
        if (t->hit) return;
        collision_recurse(t, b1, n2->first_child + 1);
    }
}

int orc_se3_collide_bvh(const orc_mesh *robot, const orc_mesh *env, const double state[7],
                        orc_counters *ctr) {
    if (robot->ntris == 0 || env->ntris == 0) {
        if (ctr) ctr->checks++;
        return 0;
    }
    trav t;
    t.m1 = robot; t.m2 = env; t.hit = 0; t.bv = 0; t.tri = 0;
    state_to_rel(state, &t.X);
    collision_recurse(&t, 0, 0);
    if (ctr) { ctr->checks++; ctr->bv_tests += t.bv; ctr->tri_tests += t.tri; }
    return t.hit;
}

static int collide_mode(const orc_mesh *robot, const orc_mesh *env, const double s[7], int mode,
                        orc_counters *ctr) {
    if (mode == 0) {
        if (ctr) ctr->checks++;
        return orc_se3_collide_brute(robot, env, s);
    }
    return orc_se3_collide_bvh(robot, env, s, ctr);
}

static int pick_threads(int threads) {
#ifdef _OPENMP
    return threads > 0 ? threads : omp_get_max_threads();
#else
    (void)threads; return 1;
#endif
}
int orc_omp_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_se3_check_states(const orc_mesh *robot, const orc_mesh *env, const double *states,
                          size_t n, uint8_t *valid, int mode, int threads, orc_counters *ctr) {
    uint64_t c = 0, bv = 0, tri = 0;
    int nt = pick_threads(threads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt) reduction(+ : c, bv, tri)
    for (long long i = 0; i < (long long)n; ++i) {
        orc_counters local = {0, 0, 0};
        valid[i] = (uint8_t)!collide_mode(robot, env, states + 7 * i, mode, &local);
        c += local.checks; bv += local.bv_tests; tri += local.tri_tests;
    }
    if (ctr) { ctr->checks += c; ctr->bv_tests += bv; ctr->tri_tests += tri; }
}

/* ------------------------------------------------------------------------------------------ */
/* nigh SE3Space<S,a,b> distance [EXT]: a*acos(|q.q'|) + b*||t-t'||                            */

static double so3_distance(const double a[4], const double b[4]) {
    double d = fabs(((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) + a[3] * b[3]);
    return d < 1.0 ? acos(d) : 0.0;
}
double orc_se3_distance(const double a[7], const double b[7], double so3w, double l2w) {
    double dx = a[4] - b[4], dy = a[5] - b[5], dz = a[6] - b[6];
    return so3w * so3_distance(a, b) + l2w * sqrt((dx * dx + dy * dy) + dz * dz);
}

/* interpolate.hpp:18-36 -> Eigen QuaternionBase::slerp [EXT] and a + (b-a)*t */
void orc_se3_interpolate(const double from[7], const double to[7], double t, double out[7]) {
    double d = ((from[0] * to[0] + from[1] * to[1]) + from[2] * to[2]) + from[3] * to[3];
    double absD = fabs(d), s0, s1;
    if (absD >= 1.0 - DBL_EPSILON) {
        s0 = 1.0 - t; s1 = t;
    } else {
        double theta = acos(absD), sinTheta = sin(theta);
        s0 = sin((1.0 - t) * theta) / sinTheta;
        s1 = sin(t * theta) / sinTheta;
    }
    if (d < 0) s1 = -s1;
    for (int i = 0; i < 4; ++i) out[i] = s0 * from[i] + s1 * to[i];
    for (int i = 4; i < 7; ++i) out[i] = from[i] + (to[i] - from[i]) * t;
}

uint64_t orc_se3_edge_steps(const double from[7], const double to[7], double so3w, double l2w,
                            double inv_step) {
    return (uint64_t)ceil(orc_se3_distance(from, to, so3w, l2w) * inv_step);
}

/* se3_rigid_body_scenario.hpp:134-178, statement for statement */
int orc_se3_check_edge(const orc_mesh *robot, const orc_mesh *env, const double from[7],
                       const double to[7], double so3w, double l2w, double inv_step, int mode,
                       orc_counters *ctr) {
    double q[7];
    if (collide_mode(robot, env, to, mode, ctr)) return 0;
    size_t steps = (size_t)ceil(orc_se3_distance(from, to, so3w, l2w) * inv_step);
    if (steps < 2) return 1;
    double delta = 1 / (double)steps;
    enum { QN = 1024 };
    size_t qmin[QN], qmax[QN];
    qmin[0] = 1; qmax[0] = steps - 1;
    size_t qStart = 0, qEnd = 1;
    while (qStart != qEnd) {
        size_t mn = qmin[qStart % QN], mx = qmax[qStart % QN];
        qStart++;
        if (mn == mx) {
            orc_se3_interpolate(from, to, mn * delta, q);
            if (collide_mode(robot, env, q, mode, ctr)) return 0;
        } else if (qEnd + 2 < qStart + QN) {
            size_t mid = (mn + mx) / 2;
            orc_se3_interpolate(from, to, mid * delta, q);
            if (collide_mode(robot, env, q, mode, ctr)) return 0;
            if (mn < mid) { qmin[qEnd % QN] = mn; qmax[qEnd % QN] = mid - 1; qEnd++; }
            if (mid < mx) { qmin[qEnd % QN] = mid + 1; qmax[qEnd % QN] = mx; qEnd++; }
        } else {
            for (size_t i = mn; i <= mx; ++i) {
                orc_se3_interpolate(from, to, i * delta, q);
                if (collide_mode(robot, env, q, mode, ctr)) return 0;
            }
        }
    }
    return 1;
}

void orc_se3_check_edges(const orc_mesh *robot, const orc_mesh *env, const double *from,
                         const double *to, size_t n, double so3w, double l2w, double inv_step,
                         uint8_t *valid, int mode, int threads, orc_counters *ctr) {
    uint64_t c = 0, bv = 0, tri = 0;
    int nt = pick_threads(threads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt) reduction(+ : c, bv, tri)
    for (long long i = 0; i < (long long)n; ++i) {
        orc_counters local = {0, 0, 0};
        valid[i] = (uint8_t)orc_se3_check_edge(robot, env, from + 7 * i, to + 7 * i, so3w, l2w,
                                               inv_step, mode, &local);
        c += local.checks; bv += local.bv_tests; tri += local.tri_tests;
    }
    if (ctr) { ctr->checks += c; ctr->bv_tests += bv; ctr->tri_tests += tri; }
}

/* ------------------------------------------------------------------------------------------ */
/* nearest neighbour: exact brute force in the nigh metrics, ties to the lowest index          */

static double nn_dist(const double *a, const double *b, int dim, int metric, double so3w,
                      double l2w) {
    if (metric == 0) return orc_se3_distance(a, b, so3w, l2w);
    double s = 0;
    for (int i = 0; i < dim; ++i) s += fabs(a[i] - b[i]);
    return s;
}

void orc_nn_nearest(const double *keys, size_t n, int dim, int metric, double so3w, double l2w,
                    const double *queries, size_t nq, int64_t *idx, double *dist, int threads) {
    int nt = pick_threads(threads);
    (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (long long qi = 0; qi < (long long)nq; ++qi) {
        double best = INFINITY;
        int64_t bi = -1;
        for (size_t j = 0; j < n; ++j) {
            double d = nn_dist(keys + j * dim, queries + qi * dim, dim, metric, so3w, l2w);
            if (d < best) { best = d; bi = (int64_t)j; }
        }
        idx[qi] = bi;
        dist[qi] = best;
    }
}

void orc_nn_knearest(const double *keys, size_t n, int dim, int metric, double so3w, double l2w,
                     const double *queries, size_t nq, int k, int64_t *idx, double *dist,
                     int threads) {
    int nt = pick_threads(threads);
    (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (long long qi = 0; qi < (long long)nq; ++qi) {
        int64_t *oi = idx + (size_t)qi * k;
        double *od = dist + (size_t)qi * k;
        int cnt = 0;
        for (size_t j = 0; j < n; ++j) {
            double d = nn_dist(keys + j * dim, queries + qi * dim, dim, metric, so3w, l2w);
            if (cnt == k && !(d < od[k - 1])) continue;
            int pos = cnt < k ? cnt : k - 1;
            /* insertion keeps (dist, idx) ascending; equal distances stay in index order */
            while (pos > 0 && od[pos - 1] > d) { od[pos] = od[pos - 1]; oi[pos] = oi[pos - 1]; --pos; }
            od[pos] = d; oi[pos] = (int64_t)j;
            if (cnt < k) ++cnt;
        }
        for (int r = cnt; r < k; ++r) { od[r] = INFINITY; oi[r] = -1; }
    }
}

#include "mplb_oracle_fetch.inc"
