/*
 * mplb.h -- C ABI of the B200-native validity-check library (libmplb.so).
 *
 * This is the drop-in boundary for the reference's scenario validity interface
 * (BerkeleyAutomation/mplambda, paths relative to the reference root):
 *
 *   include/mpl/demo/se3_rigid_body_scenario.hpp:51-64    ctor (mesh load, BVH build)   -> mplb_scene_create_se3*
 *   include/mpl/demo/se3_rigid_body_scenario.hpp:119-126  isValid(const State&)         -> mplb_check_states
 *   include/mpl/demo/se3_rigid_body_scenario.hpp:134-178  isValid(from, to)             -> mplb_check_edges
 *   include/mpl/demo/load_mesh.hpp:232-293                MeshLoad<...>::load           -> mplb_mesh_load_collada
 *   include/mpl/demo/fetch_scenario.hpp:48-63,105-173     FetchScenario ctor / isValid  -> mplb_scene_create_fetch, mplb_check_states/_edges
 *   include/mpl/prrt.hpp:61,66-68 / pcforest.hpp:80-87,107  nigh insert/nearest/nearest(k) -> mplb_nn_*
 *
 * Plain pointers and sizes only; no C++/torch types.  Every function returns MPLB_OK (0) or a
 * negative error code, and mplb_last_error() holds a thread-local message.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with MPLB_ERR_CUDA.
 *
 * Threading: a scene is immutable after creation; mplb_check_* may be called concurrently from
 * any number of host threads on the same scene (the reference calls isValid from OpenMP planner
 * threads, prrt.hpp:140-152) -- each call borrows a private stream + staging context.
 */
#ifndef MPLB_H
#define MPLB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPLB_OK 0
#define MPLB_ERR_INVALID (-1) /* bad argument (the C++ wrapper throws std::invalid_argument) */
#define MPLB_ERR_IO (-2)      /* mesh file missing / unparsable (load_mesh.hpp:250-254)        */
#define MPLB_ERR_CUDA (-3)    /* no device, CUDA failure, or kernel overflow flag               */
#define MPLB_ERR_NOMEM (-4)

typedef struct mplb_scene mplb_scene;
typedef struct mplb_nn mplb_nn;

const char *mplb_last_error(void);
/* number of visible CUDA devices (0 when none / no driver) */
int mplb_device_count(void);

/* ---------------------------------------------------------------------------------------- */
/* Mesh import: replaces MeshLoad<Mesh>::load(name, shiftToCenter, identityRootTransform)
 * (load_mesh.hpp:232-293).  Emits the triangle soup visitTriangles (load_mesh.hpp:50-89) feeds
 * to addTriangle: float64 [n][3][3].  *tris is malloc'd; release with mplb_free.  center[3]
 * (optional) receives the subtracted centroid (zero when shift_to_center == 0). */
int mplb_mesh_load_collada(const char *path, int shift_to_center, int identity_root_transform,
                           double **tris, size_t *ntris, double center[3]);
void mplb_free(void *p);

/* On-disk cache of imported meshes and built hierarchies.  The reference re-imports the .dae files and rebuilds both BVHs
 * in every lambda it starts (load_mesh.hpp:232-293, se3_rigid_body_scenario.hpp:58-59); with a cache directory set,
 * mplb_scene_create_* look both up by content (file bytes / triangle bytes + format version) and only import / build on a
 * miss.  Entries are checksummed and replaced atomically; a stale or damaged entry is rebuilt.  dir == NULL or "" turns the
 * cache off (the default unless the environment names one in MPLB_CACHE_DIR).  Hits/misses are process-wide counts. */
int mplb_set_cache_dir(const char *dir);
void mplb_cache_stats(uint64_t *hits, uint64_t *misses);

/* ---------------------------------------------------------------------------------------- */
/* SE(3) rigid-body scene: replaces SE3RigidBodyScenario's constructor
 * (se3_rigid_body_scenario.hpp:51-64).  Triangles are the robot mesh (already shifted to its
 * centre, as load(robotMesh, true, false) does) and the environment mesh, float64 [n][3][3].
 * so3_weight/l2_weight are the template weights of nigh::SE3Space<S,50,1> (line 15-18) and
 * check_resolution the ctor argument (invStepSize_ = 1/checkResolution, line 63). */
int mplb_scene_create_se3(const double *env_tris, size_t n_env, const double *robot_tris,
                          size_t n_robot, double so3_weight, double l2_weight,
                          double check_resolution, int device, mplb_scene **out);
/* Convenience: the same from the two .dae paths, exactly like the reference ctor. */
int mplb_scene_create_se3_files(const char *env_path, const char *robot_path, double so3_weight,
                                double l2_weight, double check_resolution, int device,
                                mplb_scene **out);

/* Fetch scene: replaces FetchScenario's constructor (fetch_scenario.hpp:48-63).  env_tris is the
 * environment mesh as load(envMesh, false, true) yields it (identity root transform), env_frame the
 * row-major 3x4 [R|t] of `envFrame`; the 12 collision primitives and link origins are the constants of
 * fetch_robot.hpp:38-53,217-274.  States are the 8 joint values (fetch_robot.hpp:80-99); isValid(q) =
 * !selfCollision && !inCollisionWith(env) (fetch_scenario.hpp:105-118); edges use the unscaled L1
 * distance and linear interpolation (fetch_scenario.hpp:120-173). */
int mplb_scene_create_fetch(const double *env_tris, size_t n_env, const double env_frame[12],
                            double check_resolution, int device, mplb_scene **out);
int mplb_scene_create_fetch_file(const char *env_path, const double env_frame[12],
                                 double check_resolution, int device, mplb_scene **out);
void mplb_scene_destroy(mplb_scene *scene);

/* State layouts (host or device, float64):
 *   SE3  : [n][7] = qx,qy,qz,qw,tx,ty,tz   (std::tuple<Quaterniond,Vector3d>, Eigen coeff order)
 *   Fetch: [n][8] = torso_lift, shoulder_pan, ... wrist_roll (fetch_robot.hpp:80-99)
 * mplb_scene_state_dim returns 7 or 8. */
int mplb_scene_state_dim(const mplb_scene *scene);

/* isValid(q) for n states; valid[i] = 1 iff state i is collision free.  HOST buffers; includes
 * the H2D/D2H copies.  n == 1 is the reference's single call (se3...:119-126). */
int mplb_check_states(mplb_scene *scene, const double *states, size_t n, uint8_t *valid);

/* isValid(from, to) for n edges (se3...:134-178): valid[i] = isValid(to_i) AND every
 * interpolated state at i/steps, steps = ceil(distance(from,to)/checkResolution).  `from` is
 * assumed valid like the reference's assert.  The verdict is exactly that conjunction, but the
 * samples are not enumerated: intervals of samples that nothing can reach are certified free as a
 * whole.  states_checked (optional) receives the number of states actually evaluated on the
 * device (a small fraction of the sum of steps). */
int mplb_check_edges(mplb_scene *scene, const double *from, const double *to, size_t n,
                     uint8_t *valid, uint64_t *states_checked);

/* Device-resident variants: all pointers are device memory on the scene's device; work is
 * enqueued on `stream` (a cudaStream_t passed as void*; NULL = the default stream) and NOT
 * synchronised.  steps/offsets for edges are computed by mplb_edges_plan on the host.  The edge
 * and index-named calls share one set of per-scene scratch buffers (work counters, spill lists,
 * interval stacks): enqueue them on ONE stream per scene, or order them with events.
 * mplb_check_states_device may be issued on several streams at once (each call in flight has a
 * work cursor of its own); issuing successive batches alternately on two streams lets the SMs
 * one launch has drained start on the next while its last warps finish (2^20 Alpha 1.5 poses:
 * 0.74 -> 0.69 ms per batch).  The host-buffer calls above borrow private scratch and may run
 * concurrently from any number of threads. */
int mplb_check_states_device(mplb_scene *scene, const double *d_states, size_t n,
                             uint8_t *d_valid, void *stream);
/* Host-side plan of an edge batch: steps[i] = ceil(distance*invStep) exactly as
 * se3...:139 (host libm, bit-identical to the CPU path). */
int mplb_edges_plan(mplb_scene *scene, const double *from, const double *to, size_t n,
                    uint32_t *steps);
/* (max_steps is accepted for compatibility and ignored.  SE(3) scenes only.) */
int mplb_check_edges_device(mplb_scene *scene, const double *d_from, const double *d_to,
                            const uint32_t *d_steps, size_t n, uint32_t max_steps,
                            uint8_t *d_valid, void *stream);

/* Space::distance (nigh SE3Space / L1Space) on the host, used by the wrappers' isGoal. */
double mplb_distance(const mplb_scene *scene, const double *a, const double *b);

typedef struct {
    uint64_t checks;      /* states evaluated (isValid(q) equivalents, the reference's calls_) */
    uint64_t bv_tests;    /* OBB-OBB tests executed on the device                              */
    uint64_t tri_tests;   /* FP32 filtered triangle-triangle tests                             */
    uint64_t exact_tests; /* FP64 exact re-evaluations of uncertain triangle pairs             */
    uint64_t launches;    /* kernels launched                                                  */
    uint64_t env_tris, robot_tris, env_nodes, robot_nodes;
    uint64_t stream_stalls;   /* large host batches whose input did not arrive under the running kernel and were
                                 repeated copy-then-kernel (a profiler serialising launches); 0 in normal operation */
    uint64_t streaming_off;   /* 1 while the scene sends large host batches copy-then-kernel because of such stalls */
    uint64_t ambiguous_steps; /* index-named edges whose device-planned step count was within rounding of an integer
                                 and was re-planned with host libm (mplb_check_edges_indexed)                          */
    uint64_t coalesced_calls; /* scalar calls that shared a launch with other threads' calls (see mplb_scene_set_coalescing) */
    double min_exact_margin;  /* with mplb_scene_track_margins on: the smallest relative margin |G| any exact (FP64)
                                 triangle-pair decision had since the last reset -- G = the largest gap an axis leaves
                                 between the projections over |axis|_1 x the largest coordinate; the pair is disjoint iff
                                 G > 0.  +inf: none recorded.  The device's sin/acos differ from glibc's by <= 2 ulp (edge
                                 samples): a verdict could only depend on that if this were ~1e-15 or less.               */
} mplb_stats_t;
int mplb_scene_stats(mplb_scene *scene, mplb_stats_t *out, int reset);
/* Diagnostic: the exact path evaluates all 17 axes of every uncertain pair (no early exit) and records the decision's
 * margin (mplb_stats_t.min_exact_margin); verdicts are unchanged.  SE(3) scenes; off by default (costs ~2x in the exact
 * path, which is a few per cent of a batch). */
int mplb_scene_track_margins(mplb_scene *scene, int on);

/* Scalar calls (n == 1: the reference's isValid(q) / isValid(from, to), issued one per OpenMP planner thread,
 * prrt.hpp:140-152, 280-299) go to the device as ONE kernel launch each: the kernel reads the state from pinned host
 * memory, writes the verdict there and rings the calling thread, which spins on it (no copy engine, no stream
 * synchronisation; 24-30 us per isValid(q), 16 threads: 400 k calls/s).  With coalescing on, scalar calls that arrive
 * while earlier ones of the same scene are on the device share a launch (up to 64); a lone caller is never delayed.
 * Off by default (MPLB_COALESCE=1 in the environment: on by default): with every thread on its own stream sharing
 * only pays once there are more calling threads than the driver launches side by side.  Returns MPLB_OK. */
int mplb_scene_set_coalescing(mplb_scene *scene, int on);

/* ---------------------------------------------------------------------------------------- */
/* Nearest neighbour: replaces nigh::Nigh<Node*,Space,NodeKey,Concurrent,...> as used by the
 * planners (prrt.hpp:42,61,67,115; pcforest.hpp:51,81,86,107,193).  Keys are the scaled states
 * (Scenario::scale).  metric 0 = SE3Space<double,so3w,l2w> on 7-double keys, metric 1 =
 * L1Space<double,dim>.  Indices are insertion order; ties go to the lowest index. */
int mplb_nn_create(int metric, int dim, double so3_weight, double l2_weight, int device,
                   mplb_nn **out);
void mplb_nn_destroy(mplb_nn *nn);
int mplb_nn_insert(mplb_nn *nn, const double *keys, size_t n); /* host keys [n][dim] */
size_t mplb_nn_size(const mplb_nn *nn);
/* nearest(q): idx[i] = -1 and dist = +inf when the structure is empty */
int mplb_nn_nearest(mplb_nn *nn, const double *queries, size_t nq, int64_t *idx, double *dist);
/* nearest(nbh, q, k): [nq][k] ascending by (distance, index), padded with -1 / +inf */
int mplb_nn_knearest(mplb_nn *nn, const double *queries, size_t nq, int k, int64_t *idx,
                     double *dist);
/* Beyond brute force (nigh's KD-tree strategy, prrt.hpp:35-42): from 256 keys on the searches go through a
 * uniform grid over the SE(3) keys' translations (the metric's translation term bounds the distance from below) or
 * the first three coordinates of L1 keys (a partial sum of the L1 distance), rebuilt whenever keys have come in
 * (structures of up to 65 536 keys: ~50 us) or when the keys inserted since outgrow a quarter of it (larger ones); those
 * are searched by brute force meanwhile and the lists
 * merged.  Results are the brute-force ones bit for bit.  mplb_nn_set_index(nn, 0) keeps a structure on brute force;
 * mplb_nn_index_info reports the keys indexed, the grid's cells, the FP64 distance evaluations of the indexed searches
 * so far and the queries that needed the slow exact selection (more candidates than a warp's list holds: duplicates). */
int mplb_nn_set_index(mplb_nn *nn, int on);
int mplb_nn_index_info(mplb_nn *nn, size_t *indexed, size_t *cells, uint64_t *exact_evaluations,
                       uint64_t *slow_queries);
/* device-resident 1-NN: queries/idx/dist are device pointers, enqueued on `stream`, not synchronised;
 * d_dist holds the device-evaluated distance (CUDA acos: may differ from the host value in the last bit) */
int mplb_nn_nearest_device(mplb_nn *nn, const double *d_queries, size_t nq, int64_t *d_idx,
                           double *d_dist, void *stream);

/* k-NN on device-resident queries: d_idx / d_dist are [nq][k] device arrays (layout and order as mplb_nn_knearest;
 * distances device-evaluated).  The *_device searches of one structure share its scan scratch; they are ordered among
 * themselves by an event the structure keeps, whatever streams they are enqueued on. */
int mplb_nn_knearest_device(mplb_nn *nn, const double *d_queries, size_t nq, int k, int64_t *d_idx,
                            double *d_dist, void *stream);
/* Keys that already sit in device memory (new tree nodes picked on the device); returns when they are in. */
int mplb_nn_insert_device(mplb_nn *nn, const double *d_keys, size_t n, void *stream);
/* The structure's key array in device memory, [size][dim] float64 in insertion order: the pool index-named edges
 * (below) read tree nodes from.  The pointer changes when the array grows past mplb_nn_capacity: reserve first. */
int mplb_nn_reserve(mplb_nn *nn, size_t capacity);
size_t mplb_nn_capacity(const mplb_nn *nn);
const double *mplb_nn_keys_device(const mplb_nn *nn);

/* ---------------------------------------------------------------------------------------- */
/* Device-resident planner iteration: sample -> nearest -> isValid(from, to) -> insert with only indices and counts
 * crossing PCIe (prrt.hpp:280-299; pcforest.hpp:507-567 issues the same motions for every neighbour).
 *
 * Scenario::randomSample (se3...:107-113, randomize.hpp:12-38; Fetch: fetch_robot.hpp:163-175) for n states written
 * to device memory: state i is a pure function of (seed, first + i) -- Philox4x32-10, the generator of
 * mplambda_b200/workload.py.  min/max: 3 translation bounds (SE3) or 8 joint bounds (Fetch). */
int mplb_sample_states_device(mplb_scene *scene, const double *min, const double *max, uint64_t seed,
                              uint64_t first, size_t n, double *d_states, void *stream);

/* isValid(from, to) for n edges named by index: edge e runs from row (from_idx ? from_idx[e] : e / from_div) of
 * d_from_pool to row (to_idx ? to_idx[e] : e / to_div) of d_to_pool (device arrays [rows][7], e.g.
 * mplb_nn_keys_device: the tree's nodes).  An index outside its pool makes the edge invalid.  steps =
 * ceil(distance/resolution) is planned on the device; where the device's acos could round it differently from the
 * host's, the edge is planned and checked again through the host path, so verdicts equal mplb_check_edges' on the
 * same states (mplb_stats_t.ambiguous_steps counts those edges).  The index arrays are host memory (copied in: 8 B
 * per end point instead of 56) or, with idx_on_device != 0, device memory; `valid` may be host or device memory.
 * Synchronous.  SE(3) scenes. */
int mplb_check_edges_indexed(mplb_scene *scene, const double *d_from_pool, size_t from_rows,
                             const int64_t *from_idx, uint32_t from_div, const double *d_to_pool,
                             size_t to_rows, const int64_t *to_idx, uint32_t to_div, size_t n,
                             int idx_on_device, uint8_t *valid, uint64_t *states_checked);
/* The same enqueued on `stream` without synchronisation, every pointer device memory.  Ambiguous step counts are only
 * counted (mplb_stats_t.ambiguous_steps after mplb_scene_stats), not re-planned; shares the per-scene scratch of the
 * other *_device calls (one stream per scene). */
int mplb_check_edges_indexed_device(mplb_scene *scene, const double *d_from_pool, size_t from_rows,
                                    const int64_t *d_from_idx, uint32_t from_div,
                                    const double *d_to_pool, size_t to_rows, const int64_t *d_to_idx,
                                    uint32_t to_div, size_t n, uint8_t *d_valid, void *stream);

/* Planner<Scenario, PRRT>'s loop (prrt.hpp:280-299) run on the device, `batch` samples per iteration: the tree's
 * nodes are the keys of `nn` (which must be empty, on the scene's device, SE3 metric with the scene's weights), node
 * i's parent is parent[i].  Per iteration the host reads a 24-byte status.  Samples of one batch do not see each
 * other (like the reference's concurrent threads); accepted samples become nodes in sample order, so the tree is a
 * pure function of (scene, start, goal, bounds, seed, batch, goal_bias).  Sample s of iteration t is draw t * batch + s
 * of mplb_sample_states_device's stream, replaced by `goal` when its seventh uniform is < goal_bias
 * (prrt.hpp:301-315).  Fails with MPLB_ERR_INVALID when `start` is not valid (prrt.hpp:109-110). */
typedef struct mplb_rrt mplb_rrt;
typedef struct {
    uint64_t iterations, samples, edges_checked, nodes;
    int64_t solution; /* first goal node (-1: none yet) */
    double seconds;   /* wall time of the last mplb_rrt_run */
} mplb_rrt_status_t;
int mplb_rrt_create(mplb_scene *scene, mplb_nn *nn, const double *start, const double *goal,
                    double goal_radius, const double *min, const double *max, uint64_t seed,
                    size_t batch, double goal_bias, mplb_rrt **out);
void mplb_rrt_destroy(mplb_rrt *rrt);
/* runs until a goal node exists, max_iterations more iterations have run or time_limit_s (> 0) has passed */
int mplb_rrt_run(mplb_rrt *rrt, uint64_t max_iterations, double time_limit_s, mplb_rrt_status_t *status);
/* the tree: nodes [count][7] and parent [count] (root: -1), at most `capacity` of them; *count = nodes there are */
int mplb_rrt_tree(mplb_rrt *rrt, double *nodes, int64_t *parent, size_t capacity, size_t *count);
/* the sample batch of iteration `iteration` as the loop draws it ([batch][7], [batch] goal-sample flags): lets a CPU
 * planner replay the same inputs (tests) */
int mplb_rrt_samples(mplb_rrt *rrt, uint64_t iteration, double *states, uint8_t *is_goal_sample);

#ifdef __cplusplus
}
#endif
#endif /* MPLB_H */
