/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed
 * by the product path (mplambda_b200/, include/mplb*.h).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and there only as the
 * checker / the reported CPU baseline.
 *
 * FP64 CPU restatement of the arithmetic behind the reference's scenario validity interface
 * (/root/reference/include/mpl/demo/se3_rigid_body_scenario.hpp:119-178,
 *  fetch_scenario.hpp:105-173, fetch_robot.hpp:276-315,502-647) and of the nearest-neighbour
 * query that precedes it (include/mpl/prrt.hpp:66-68, pcforest.hpp:80-87).
 *
 * PARITY UNPINNED: the arithmetic itself lives in third-party libraries that are NOT vendored
 * in /root/reference and not installed in this image -- FCL (>=0.6, git master, unpinned,
 * CMakeLists.txt:16), libccd (>=2.0, unpinned, CMakeLists.txt:17), nigh (sibling checkout,
 * unversioned, CMakeLists.txt:71), Eigen (>=3.3) and assimp.  The reference's own tests
 * (test/fetch_robot_{jacobian,ik}_test.cpp) hold no golden vector for any function on this
 * path.  What follows restates the published algorithms of those libraries (marked [EXT]);
 * it is anchored on the reference's call sites and on the weak fixtures the repository does
 * hold (scripted start/goal states must be valid, scripts/Twistycool.{2,3}.path must be valid
 * paths) -- see tests/test_oracle_fixtures.py.
 *
 * Compile with -ffp-contract=off: the operation order below is the specification the CUDA
 * exact path mirrors.
 */
#ifndef MPLB_ORACLE_H
#define MPLB_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_mesh orc_mesh;

/* tris: [n][3][3] doubles (the triangle soup load_mesh.hpp:275-284 feeds to addTriangle). */
orc_mesh *orc_mesh_create(const double *tris, size_t n);
void orc_mesh_destroy(orc_mesh *m);
size_t orc_mesh_num_tris(const orc_mesh *m);
size_t orc_mesh_num_nodes(const orc_mesh *m);

typedef struct {
    uint64_t checks;   /* isValid(q) evaluations             */
    uint64_t bv_tests; /* OBB-OBB tests in FCL traversal order */
    uint64_t tri_tests;/* triangle-triangle tests             */
} orc_counters;

/* state = qx,qy,qz,qw,tx,ty,tz (Eigen coeff order, se3_rigid_body_scenario.hpp:18-19). */

/* O1: brute-force all-pairs 17-axis SAT.  returns 1 if robot(pose) collides with env. */
int orc_se3_collide_brute(const orc_mesh *robot, const orc_mesh *env, const double state[7]);
/* O2: FCL traversal order (collisionRecurse, first-hit exit). */
int orc_se3_collide_bvh(const orc_mesh *robot, const orc_mesh *env, const double state[7],
                        orc_counters *ctr);

/* isValid(q) over a batch; mode 0 = O1 brute, 1 = O2 bvh.  threads<=0 -> omp max. */
void orc_se3_check_states(const orc_mesh *robot, const orc_mesh *env, const double *states,
                          size_t n, uint8_t *valid, int mode, int threads, orc_counters *ctr);

double orc_se3_distance(const double a[7], const double b[7], double so3w, double l2w);
void orc_se3_interpolate(const double from[7], const double to[7], double t, double out[7]);
/* number of interpolation steps of isValid(from,to) (se3_rigid_body_scenario.hpp:139) */
uint64_t orc_se3_edge_steps(const double from[7], const double to[7], double so3w, double l2w,
                            double inv_step);
/* isValid(from,to) in the reference's bisection order with first-failure return
 * (se3_rigid_body_scenario.hpp:134-178).  states_checked counts isValid(q) calls made. */
int orc_se3_check_edge(const orc_mesh *robot, const orc_mesh *env, const double from[7],
                       const double to[7], double so3w, double l2w, double inv_step, int mode,
                       orc_counters *ctr);
void orc_se3_check_edges(const orc_mesh *robot, const orc_mesh *env, const double *from,
                         const double *to, size_t n, double so3w, double l2w, double inv_step,
                         uint8_t *valid, int mode, int threads, orc_counters *ctr);

/* Nearest neighbour (nigh [EXT]): metric 0 = SE3(so3w,l2w) on 7-double keys,
 * metric 1 = L1 on dim-double keys.  Exact argmin, ties to the lowest index. */
void orc_nn_nearest(const double *keys, size_t n, int dim, int metric, double so3w, double l2w,
                    const double *queries, size_t nq, int64_t *idx, double *dist, int threads);
/* k nearest, ascending by (distance, index).  idx/dist are [nq][k]; short lists padded -1/inf */
void orc_nn_knearest(const double *keys, size_t n, int dim, int metric, double so3w, double l2w,
                     const double *queries, size_t nq, int k, int64_t *idx, double *dist,
                     int threads);

int orc_omp_max_threads(void);

/* ---- Fetch (fetch_robot.hpp / fetch_scenario.hpp); state = 8 joint values (fetch_robot.hpp:80-99) ---- */
/* 12 geometry frames (row-major 3x4 [R|t]) in the order of fetch_robot.hpp:38-53, and gripperAxis z */
void orc_fetch_fk(const double q[8], double frames[12][12], double *gripper_axis_z);
/* 0 none, 1..24 colliding pair (order of fetch_robot.hpp:508-577), 25 floor clearance (:580) */
int orc_fetch_self_collision(const double q[8]);
int orc_fetch_is_valid(const orc_mesh *env, const double env_frame[12], const double q[8], int *why,
                       orc_counters *ctr);
void orc_fetch_check_states(const orc_mesh *env, const double env_frame[12], const double *states, size_t n,
                            uint8_t *valid, int threads, orc_counters *ctr);
double orc_fetch_distance(const double a[8], const double b[8]);
uint64_t orc_fetch_edge_steps(const double a[8], const double b[8], double inv_step);
int orc_fetch_check_edge(const orc_mesh *env, const double env_frame[12], const double from[8], const double to[8],
                         double inv_step, orc_counters *ctr);
void orc_fetch_check_edges(const orc_mesh *env, const double env_frame[12], const double *from, const double *to,
                           size_t n, double inv_step, uint8_t *valid, int threads, orc_counters *ctr);

#ifdef __cplusplus
}
#endif
#endif
