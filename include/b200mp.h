/* b200mp -- B200-native batched motion validation and nearest-neighbour engine: C ABI.
 *
 * This is the drop-in boundary. Host code (the OMPL adapters in include/b200mp_ompl.hpp, or any FFI) talks to
 * the CUDA engine only through these entry points: opaque handles, plain pointers and sizes, int status codes,
 * no exceptions and no C++/torch types cross it. Paths below are relative to the reference checkout.
 *
 * What each entry point replaces in the reference:
 *   b200mp_validate_states    ompl::vamp::VampStateValidityChecker::isValid      src/ompl/vamp/VampStateValidityChecker.h:37-41
 *                             (= vamp::planning::validate_motion<Robot, 8, 1>(q, q, env))
 *   b200mp_validate_motions   ompl::vamp::VampMotionValidator::checkMotion       src/ompl/vamp/VampMotionValidator.h:37-52
 *                             (= validate_motion<Robot, 8, Robot::resolution>)  external/vamp/src/impl/vamp/planning/validate.hh:23-77
 *   b200mp_env_*              vamp::collision::Environment<float> + sort()       external/vamp/src/impl/vamp/collision/environment.hh:14-85
 *                             and the shape constructors' min_distance           external/vamp/src/impl/vamp/collision/shapes.hh:52-67,165-189,236-239
 *   b200mp_env_set_attachment vamp::collision::Attachment, Robot::fkcc_attach   external/vamp/src/impl/vamp/collision/attachments.hh:13-56,
 *                                                                                external/vamp/src/impl/vamp/robots/panda.hh:10264-15469
 *   b200mp_env_add_capt       vamp::collision::CAPT (prebuilt tree arrays)       external/vamp/src/impl/vamp/collision/capt.hh:61-618
 *   b200mp_env_add_pointcloud vamp::collision::CAPT::CAPT / subdivide            external/vamp/src/impl/vamp/collision/capt.hh:104-365
 *   b200mp_robot_info         Robot::dimension / resolution / joint bounds       src/ompl/vamp/VampStateSpace.h:10-36
 *   b200mp_chain_*            KinematicChainSpace::distance / interpolate,        demos/KinematicChain.h:105-121,150-176
 *                             KinematicChainValidityChecker::isValid,             demos/KinematicChain.h:209-276
 *                             createHornEnvironment,                              demos/KinematicChain.h:279-315
 *                             DiscreteMotionValidator::checkMotion (both forms)   src/ompl/base/src/DiscreteMotionValidator.cpp:48-145
 *   b200mp_knn_*              ompl::NearestNeighbors<_T> (Linear semantics)      src/ompl/datastructures/NearestNeighbors.h:46-115,
 *                                                                                src/ompl/datastructures/NearestNeighborsLinear.h:65-142
 *                             metric: RealVectorStateSpace::distance            src/ompl/base/spaces/src/RealVectorStateSpace.cpp:230-242
 *
 * Conventions: every function returns B200MP_OK (0) or an error code; b200mp_last_error() gives the text for the
 * calling thread. Batch buffers are caller-owned, contiguous, row-major. "_device" variants take device pointers
 * that live on the context's GPU and an optional cudaStream_t passed as void* (NULL = the context's own stream);
 * they do not synchronise. Host variants are synchronous and thread-safe: small validation calls (a planner thread's single
 * isValid / checkMotion) run concurrently on per-call lanes (own stream + mapped buffer), large ones are serialised per context.
 * There is no CPU fallback: without a CUDA device b200mp_create fails with B200MP_ERR_NO_DEVICE.
 */
#ifndef B200MP_H
#define B200MP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MP_VERSION 2

typedef struct b200mp_ctx b200mp_ctx;
typedef struct b200mp_env b200mp_env;
typedef struct b200mp_knn b200mp_knn;
typedef struct b200mp_capt b200mp_capt;
typedef struct b200mp_chain b200mp_chain;
typedef struct b200mp_comm b200mp_comm;
typedef struct b200mp_session b200mp_session;

enum {
    B200MP_OK = 0,
    B200MP_ERR_INVALID = 1,    /* bad argument */
    B200MP_ERR_CUDA = 2,       /* a CUDA runtime call failed; see b200mp_last_error() */
    B200MP_ERR_NO_DEVICE = 3,  /* no usable GPU (there is no CPU fallback) */
    B200MP_ERR_CAPACITY = 4,   /* scene or request exceeds an engine limit */
    B200MP_ERR_STATE = 5,      /* call order violated, e.g. environment not committed */
    B200MP_ERR_EMPTY = 6       /* "No elements found in nearest neighbors data structure" */
};

enum { B200MP_ROBOT_PANDA = 0, B200MP_ROBOT_FETCH = 1, B200MP_ROBOT_UR5 = 2, B200MP_ROBOT_BAXTER = 3 };
enum { B200MP_METRIC_L2 = 0, B200MP_METRIC_CHAIN = 1 };

int b200mp_version(void);
const char *b200mp_last_error(void);

/* ---- context: one per GPU (several GPUs: one context each, grouped by b200mp_comm_*, see below) ----------------- */
int b200mp_create(int device, b200mp_ctx **out);
void b200mp_destroy(b200mp_ctx *ctx);
int b200mp_device(const b200mp_ctx *ctx, int *device, int *sm_count);
int b200mp_synchronize(b200mp_ctx *ctx);

/* ---- robots ----------------------------------------------------------------------------------------------- */
/* lo/hi: dim floats each (may be NULL). */
int b200mp_robot_info(int robot, int *dim, int *n_spheres, int *resolution, float *lo, float *hi);

/* ---- environment -------------------------------------------------------------------------------------------- */
int b200mp_env_create(b200mp_ctx *ctx, b200mp_env **out);
void b200mp_env_destroy(b200mp_env *env);
/* xyzr[n][4] */
int b200mp_env_add_spheres(b200mp_env *env, const float *xyzr, int n);
/* c[n][15] = centre(3), axis1(3), axis2(3), axis3(3), half extents(3). z_aligned selects the reference's
 * z_aligned_cuboids list (only axis1.xy, axis2.xy and half extents are then used). */
int b200mp_env_add_cuboids(b200mp_env *env, const float *c, int n, int z_aligned);
/* c[n][7] = p1(3), v = p2 - p1 (3), radius. rdv = 1 / (v.v) is derived here. */
int b200mp_env_add_capsules(b200mp_env *env, const float *c, int n, int z_aligned);
/* Prebuilt collision-affording point tree, arrays laid out as the reference stores them:
 * tests[2^nlog2 - 1], aabbs[2^nlog2][6] (lower xyz, upper xyz), aff_starts[2^nlog2 + 1] (in units of 8-point
 * vectors), aff[3][n_aff8 * 8] (x row, y row, z row), top[6]. */
int b200mp_env_add_capt(b200mp_env *env, int nlog2, const float *tests, const float *aabbs, const uint32_t *aff_starts,
                        const float *aff, int64_t n_aff8, const float *top, float r_min, float r_max, float r_point);
/* Build the tree from raw points on the host exactly as the reference's constructor does
 * (vamp::collision::CAPT(points, r_min, r_max, r_point), collision/capt.hh:294-365) and attach it. xyz[n][3]. */
int b200mp_env_add_pointcloud(b200mp_env *env, const float *xyz, int64_t n, float r_min, float r_max, float r_point);
/* The tree built on the device instead (n <= 2^17 points, i.e. a filtered cloud): histogram-median split planes and one block
 * per leaf gathering its list. Same structure and query as the reference's tree, not the same arrays. Every leaf lists ALL
 * points a sphere of radius <= r_max centred in its cell can touch, so the query is exact (equal to brute force over the
 * points). The reference's own build can drop such a point while handing candidates down its recursion (reproduced by
 * b200mp_env_add_pointcloud, whose arrays equal the reference's), so this tree accepts no state the reference-identical tree
 * rejects, and rejects a few it accepts (0.02 % of random Fetch states in tests/test_capt_gpu.py). Use the host build when
 * decisions must equal the reference's bit for bit. */
int b200mp_env_add_pointcloud_device(b200mp_env *env, const float *xyz, int64_t n, float r_min, float r_max, float r_point);
/* The same build as a standalone host object (no GPU needed), e.g. to export the arrays.
 * sizes[4] = nlog2, n_tests, n_leaves, n_aff8. Any output pointer of b200mp_capt_export may be NULL. */
int b200mp_capt_build(const float *xyz, int64_t n, float r_min, float r_max, float r_point, b200mp_capt **out);
void b200mp_capt_destroy(b200mp_capt *capt);
int b200mp_capt_sizes(const b200mp_capt *capt, int64_t *sizes);
int b200mp_capt_export(const b200mp_capt *capt, float *tests, float *aabbs, uint32_t *aff_starts, float *aff, float *top);
/* vamp::collision::factory::heightfield::flat (collision/factory.hh:363-386; shape collision/shapes.hh:250-312, test
 * collision/sphere_heightfield.hh:9-30): centre, cell scale, nx x ny cells, heights[ny][nx] row-major (host buffer, copied).
 * Tested after the z-aligned cuboids and before the point clouds, without pruning, like the reference (validity.hh:132-138). */
int b200mp_env_add_heightfield(b200mp_env *env, float cx, float cy, float cz, float sx, float sy, float sz, int64_t nx, int64_t ny,
                               const float *heights);
/* vamp::collision::filter_pointcloud (external/vamp/src/impl/vamp/collision/filter.hh:176-274; Python entry
 * vamp.filter_pointcloud, bindings/environment.cc:168-190) on the device: cull to the range ball around `origin` and the
 * workspace box (when cull != 0), then six Morton-curve passes that drop every point within min_dist of the last point
 * kept. xyz[n][3] and out_xyz[>= n + 1][3] are host buffers; *out_n receives the number of points written (the reference
 * can emit point 0 one extra time, hence n + 1). Points with equal curve codes keep their previous relative order (the
 * reference's unstable sort leaves that order unspecified). */
int b200mp_filter_pointcloud(b200mp_ctx *ctx, const float *xyz, int64_t n, float min_dist, float max_range, const float *origin,
                             const float *ws_min, const float *ws_max, int cull, float *out_xyz, int64_t *out_n);
/* vamp::collision::Attachment + Environment::attachments (collision/attachments.hh:13-56, environment.hh:25,87-93): spheres
 * rigidly fixed to the robot's end effector (a grasped object). tf[12] = row-major 3x4 [R | t], the pose of the attachment
 * relative to the end-effector frame (Robot::eefk); xyzr[n][4] = its spheres in the attachment's own frame. While an attachment
 * is set every validation call runs Robot::fkcc_attach instead of fkcc (planning/validate.hh:43,58): the robot's own tests, then
 * the posed attachment against the environment and against the links the robot's table lists (robots/panda.hh:15308-15466).
 * n = 0 removes it. May be called before or after b200mp_env_commit (no re-commit needed); not while a validation call on this
 * environment is in flight on another thread. */
int b200mp_env_set_attachment(b200mp_env *env, const float *tf, const float *xyzr, int n);
/* Sort every shape list by min_distance (Environment::sort). Obstacles are otherwise tested in insertion
 * order with the reference's ordered-scan-with-break semantics. */
int b200mp_env_sort(b200mp_env *env);
/* Sorted scenes with >= 32 obstacles are indexed at commit by a uniform grid of candidate lists so a robot sphere is
 * tested against the few obstacles near it instead of the whole list (same decisions; see GridView in
 * ompl_b200/csrc/validate.cuh). enable = 0 keeps the plain ordered scan (A/B measurements, tests). Default: enabled. */
int b200mp_env_set_grid(b200mp_env *env, int enable);
/* Upload to the GPU. Must be called after the last add/sort and before validation. */
int b200mp_env_commit(b200mp_env *env);
/* kind: 0 spheres, 1 capsules, 2 z-capsules, 3 cuboids, 4 z-cuboids. out may be NULL. */
int64_t b200mp_env_min_distances(const b200mp_env *env, int kind, float *out);

/* ---- validation --------------------------------------------------------------------------------------------- */
/* q[m][dim] host floats -> valid[m] (1 = collision free). */
int b200mp_validate_states(b200mp_ctx *ctx, int robot, const b200mp_env *env, const float *q, int64_t m, uint8_t *valid);
/* q0, q1[m][dim] host floats -> valid[m]. resolution 0 = the robot's default (32). The states tested on an edge
 * are exactly the reference's: t = j/(8n), j = 1..8n, n = max(ceil(|q1-q0| / 8 * resolution), 1). */
int b200mp_validate_motions(b200mp_ctx *ctx, int robot, const b200mp_env *env, const float *q0, const float *q1, int64_t m,
                            int resolution, uint8_t *valid);
/* Same with double states as OMPL stores them (cast to float per ompl_to_vamp, src/ompl/vamp/Utils.h:16-30). */
int b200mp_validate_states_f64(b200mp_ctx *ctx, int robot, const b200mp_env *env, const double *q, int64_t m, uint8_t *valid);
int b200mp_validate_motions_f64(b200mp_ctx *ctx, int robot, const b200mp_env *env, const double *q0, const double *q1,
                                int64_t m, int resolution, uint8_t *valid);
/* Device-resident variants (asynchronous on `stream`). */
int b200mp_validate_states_device(b200mp_ctx *ctx, int robot, const b200mp_env *env, const float *d_q, int64_t m,
                                  uint8_t *d_valid, void *stream);
int b200mp_validate_motions_device(b200mp_ctx *ctx, int robot, const b200mp_env *env, const float *d_q0, const float *d_q1,
                                   int64_t m, int resolution, uint8_t *d_valid, void *stream);
/* Motions given as index pairs into one device-resident state table: edge e runs from d_states[d_edges[2e]] to
 * d_states[d_edges[2e+1]] (roadmap edges; PRM::addMilestone checks neighbour -> new milestone, PRM.cpp:585-598). */
int b200mp_validate_edges_device(b200mp_ctx *ctx, int robot, const b200mp_env *env, const float *d_states, const int32_t *d_edges,
                                 int64_t m, int resolution, uint8_t *d_valid, void *stream);
/* Number of kernel launches issued through this context so far (bench.py's gpu_launches). */
int64_t b200mp_launch_count(const b200mp_ctx *ctx);
/* ---- sessions: single calls without a launch ----------------------------------------------------------------
 * A serial planner (RRT-Connect, BASELINE config 1) asks one VampMotionValidator::checkMotion / VampStateValidityChecker::isValid at a
 * time (src/ompl/vamp/VampMotionValidator.h:37-52, VampStateValidityChecker.h:37-41); through b200mp_validate_motions that is one kernel
 * launch and one stream synchronise per call (~30 us). A session keeps ONE small block resident on the GPU -- the scene staged in its
 * shared memory -- polling a mailbox in mapped host memory: a call writes the two configurations and a sequence number and spins on the
 * answer. Decisions are those of b200mp_validate_motions (same states, same tests). The kernel leaves by itself after 50 ms without
 * a call (B200MP_SESSION_IDLE_MS) and is restarted by the next one, so an idle session holds no SM. A session reads the committed
 * environment's device arrays: while one is open, b200mp_env_commit and b200mp_env_set_attachment on that environment fail with
 * B200MP_ERR_STATE (close the session, change the scene, open a new one); destroy sessions before their environment. Calls on one
 * session are serialised. */
int b200mp_session_open(b200mp_ctx *ctx, int robot, const b200mp_env *env, b200mp_session **out);
void b200mp_session_close(b200mp_session *session);
/* q0, q1[dim]; resolution 0 = the robot's default. */
int b200mp_session_check_motion(b200mp_session *session, const float *q0, const float *q1, int resolution, uint8_t *valid);
int b200mp_session_check_motion_f64(b200mp_session *session, const double *q0, const double *q1, int resolution, uint8_t *valid);
int b200mp_session_is_valid(b200mp_session *session, const float *q, uint8_t *valid);
int b200mp_session_is_valid_f64(b200mp_session *session, const double *q, uint8_t *valid);
/* calls answered so far, kernel (re)starts so far; either pointer may be NULL */
int b200mp_session_stats(const b200mp_session *session, int64_t *calls, int64_t *launches);

/* Tuning switches (decisions never depend on them). B200MP_OPT_FIRST_LOOK (default 1): batches of long motions (>= 5 rakes) first get a
 * look at two states of rake 0 (the middle and the end of the motion; roadmap edges: 3/8 and 6/8), sixteen motions per warp pass,
 * continued only while it settles enough of them (see validate_rakes_kernel); 0 sends them straight to the warp-per-motion kernel
 * (measurements of that kernel alone). */
enum { B200MP_OPT_FIRST_LOOK = 1 };
int b200mp_set_option(b200mp_ctx *ctx, int option, int value);

/* ---- nearest neighbours ------------------------------------------------------------------------------------- */
/* A container of `dim`-dimensional points with the NearestNeighbors<_T> contract. Points get consecutive ids in
 * insertion order (the adapter maps ids <-> _T). Distances are the reference's doubles: sqrt(sum (a_i - b_i)^2). */
int b200mp_knn_create(b200mp_ctx *ctx, int dim, int metric, b200mp_knn **out);
/* Same container under KinematicChainSpace::distance (demos/KinematicChain.h:105-121): points are n_links joint angles. */
int b200mp_knn_create_chain(b200mp_ctx *ctx, int n_links, double link_length, b200mp_knn **out);
void b200mp_knn_destroy(b200mp_knn *knn);
int b200mp_knn_clear(b200mp_knn *knn);
/* Batches of >= 4096 queries against >= 32768 points are answered through a cell index: a balanced k-d partition of the
 * database built on the device (leaves of ~64 rows stored contiguously, with tight bounding boxes for the leaves and for
 * their 16- and 256-leaf subtrees), queries grouped by leaf, and an exact search that visits only the leaves within a
 * query's current k-th distance (lower bounds only, so the results are those of the exhaustive kernel).
 * enable = 0 forces the exhaustive tiled kernel (A/B measurements, tests). Default: enabled. */
int b200mp_knn_set_index(b200mp_knn *knn, int enable);
/* Pruning statistics of the cell-index search, summed since creation and collected only while the environment variable
 * B200MP_KNN_STAT_PAIRS is set: b200mp_knn_rows_visited = (query, candidate row) pairs evaluated (the exhaustive kernel
 * evaluates size x queries); b200mp_knn_search_stats fills out[4] = pairs evaluated, exact fp64 re-evaluations, list
 * insertions, (query batch, leaf) visits. */
int64_t b200mp_knn_rows_visited(b200mp_knn *knn);
int b200mp_knn_search_stats(b200mp_knn *knn, int64_t *out);
int64_t b200mp_knn_size(const b200mp_knn *knn);
/* pts[n][dim] host doubles; *first_id (may be NULL) receives the id of pts[0]. */
int b200mp_knn_add(b200mp_knn *knn, const double *pts, int64_t n, int64_t *first_id);
/* Append device-resident fp32 points (exactly representable in fp64); max_norm >= max |p|. Asynchronous on `stream`. */
int b200mp_knn_add_device_f32(b200mp_knn *knn, const float *d_pts, int64_t n, float max_norm, void *stream);
int b200mp_knn_remove(b200mp_knn *knn, int64_t id);
/* q[nq][dim] -> idx[nq][k], dist[nq][k] ascending by (distance, id); missing entries are -1 / +inf and counts[i]
 * (may be NULL) is the number found. self_ids (may be NULL): id that query i must not return (PRM queries a milestone
 * before inserting it, src/ompl/geometric/planners/prm/src/PRM.cpp:583,601). k <= 160. */
int b200mp_knn_query(b200mp_knn *knn, const double *q, int64_t nq, int k, const int64_t *self_ids, int64_t *idx, double *dist,
                     int *counts);
/* Device buffers. exclude_self != 0: query i skips d_self_ids[i], or id self_base + i when d_self_ids is NULL. */
int b200mp_knn_query_device(b200mp_knn *knn, const double *d_q, int64_t nq, int k, const int64_t *d_self_ids, int64_t self_base,
                            int exclude_self, int64_t *d_idx, double *d_dist, int *d_counts, void *stream);
/* Same with fp32 queries (converted exactly to fp64 on the device); query i skips id self_base + i when exclude_self. */
int b200mp_knn_query_device_f32(b200mp_knn *knn, const float *d_q, int64_t nq, int k, int64_t self_base, int exclude_self,
                                int64_t *d_idx, double *d_dist, int *d_counts, void *stream);
/* Multi-GPU query sharing: computes the rows of part `part` of `n_parts` disjoint subsets that together cover the nq
 * queries (whole index leaves per part, contiguous in k-d order and balanced by query count, when the cell index is used;
 * contiguous blocks otherwise) and leaves all other
 * rows of d_idx / d_dist untouched, so that tables pre-filled with -1 / +inf combine across ranks with an element-wise
 * max / min reduction. Every rank must hold the same database. */
int b200mp_knn_query_part_device_f32(b200mp_knn *knn, const float *d_q, int64_t nq, int k, int64_t self_base, int exclude_self, int part,
                                     int n_parts, int64_t *d_idx, double *d_dist, void *stream);
/* The same sharing for roadmap construction: only the neighbour ids, as int32, into d_ids[nq][k] (rows of other parts untouched;
 * a table pre-filled with -1 combines across GPUs with ONE element-wise max all-reduce of nq * k * 4 bytes). Always answered
 * through the cell index, whatever the batch size. k <= 64. exclude_self = 2 selects the CAUSAL query of incremental PRM: query i
 * sees only the points with ids below self_base + i (what nn_->nearestK returns when milestone i is inserted, PRM.cpp:583,601);
 * rows of early queries hold fewer than k ids, padded with -1. */
int b200mp_knn_query_part_ids_device_f32(b200mp_knn *knn, const float *d_q, int64_t nq, int k, int64_t self_base, int exclude_self, int part,
                                         int n_parts, int32_t *d_ids, void *stream);
/* nearestR: all points with distance <= radius, ascending; *count = how many exist, at most cap are written.
 * radius = +inf gives the exhaustive sorted list used for nearestK with k >= size. */
int b200mp_knn_radius(b200mp_knn *knn, const double *q, double radius, int64_t *idx, double *dist, int64_t cap, int64_t *count);

/* ---- planar kinematic chain (fp64) ------------------------------------------------------------------------------ */
/* createHornEnvironment(d, eps): *n_segments = 2 (d - 1); segments[n][4] = x0, y0, x1, y1 (may be NULL to query the count). */
int b200mp_chain_horn_environment(int d, double eps, double *segments, int *n_segments);
/* n_links <= 31 revolute joints in [-pi, pi), link length, environment segments[n_seg][4]; lvs_fraction = the state
 * space's longest valid segment fraction (OMPL default 0.01) of the maximum extent 2 pi sqrt(n). */
int b200mp_chain_create(b200mp_ctx *ctx, int n_links, double link_length, const double *segments, int n_seg, double lvs_fraction,
                        b200mp_chain **out);
void b200mp_chain_destroy(b200mp_chain *chain);
double b200mp_chain_longest_valid_segment(const b200mp_chain *chain);
/* KinematicChainSpace::distance(a[i], b[i]) and ::interpolate(a[i], b[i], t[i]); host buffers, rows of n_links doubles. */
int b200mp_chain_distance(b200mp_chain *chain, const double *a, const double *b, int64_t m, double *out);
int b200mp_chain_interpolate(b200mp_chain *chain, const double *a, const double *b, const double *t, int64_t m, double *out);
/* KinematicChainValidityChecker::isValid: no self intersection and no intersection with the environment. */
int b200mp_chain_validate_states(b200mp_chain *chain, const double *q, int64_t m, uint8_t *valid);
/* DiscreteMotionValidator::checkMotion(a[i], b[i]): valid[i]; last_valid_t[i] (may be NULL) = lastValid.second of the
 * three-argument form, -1 when the motion is valid; nd[i] (may be NULL) = validSegmentCount. */
int b200mp_chain_check_motions(b200mp_chain *chain, const double *a, const double *b, int64_t m, uint8_t *valid, double *last_valid_t, int *nd);
int b200mp_chain_validate_states_device(b200mp_chain *chain, const double *d_q, int64_t m, uint8_t *d_valid, void *stream);
int b200mp_chain_check_motions_device(b200mp_chain *chain, const double *d_a, const double *d_b, int64_t m, uint8_t *d_valid,
                                      double *d_last_valid_t, int *d_nd, void *stream);

/* ---- batched PRM roadmap construction ------------------------------------------------------------------------- */
/* Replaces the per-milestone loop of ompl::geometric::PRM (growRoadmap / addMilestone,
 * src/ompl/geometric/planners/prm/src/PRM.cpp:240-270,570-604; KStrategy, ConnectionStrategy.h:78-82) by one pass
 * per phase over the whole batch: sample -> state validity -> kNN (self excluded) -> edge list -> edge validity.
 * NOT the graph stock PRM grows from the same states: addMilestone queries only the milestones inserted EARLIER (it asks
 * before nn_->add(m), PRM.cpp:583,601), the batch connects every state to its k nearest among ALL n states and stores the
 * union as undirected edges -- a denser, insertion-order independent k-NN graph (the s-PRM construction). Edge direction
 * (older -> newer milestone) and the motion checks themselves are the reference's. */
typedef struct b200mp_roadmap_stats {
    int64_t n_states, n_candidates, n_edges, n_valid_edges;
    double ms_sample, ms_knn, ms_edges, ms_validate, ms_total; /* device time per phase on the calling rank (CUDA events) */
    double ms_gather;                                          /* final exchange of the ranks' edge lists (0 on one GPU) */
    int64_t n_local_edges;                                     /* edges this rank listed and validated */
} b200mp_roadmap_stats;
/* Phase 1. Candidates cand_first .. cand_first + cand_count - 1 of the sample stream of `seed` (candidate c is a pure
 * function of (seed, c): Philox4x32-10 -> uniform in the robot's joint box) are validated and the valid ones are
 * written, in stream order, to d_states[capacity][dim]. *n_valid = number of valid candidates (it may exceed
 * capacity; only the first `capacity` are stored). Synchronises `stream`. */
int b200mp_prm_sample_device(b200mp_ctx *ctx, int robot, const b200mp_env *env, uint64_t seed, uint64_t cand_first, int64_t cand_count,
                             float *d_states, int64_t capacity, int64_t *n_valid, void *stream);
/* Phase 3. From the neighbour table d_nbr[n][k] of ALL states (ids, -1 = none) keep, for the queries
 * q_first .. q_first + q_count - 1, edge (i, j = d_nbr[i][s]) when j does not list i, or when it does and i is the pair's
 * designated keeper (the lower index if i + j is even, the higher if odd); written in (i, s)
 * order as (min, max) index pairs to d_edges[capacity][2], with d_weights[e] = d_dist_local[(i - q_first) * k + s]
 * (both may be NULL). Synchronises `stream`. */
int b200mp_prm_edges_device(b200mp_ctx *ctx, const int64_t *d_nbr, const double *d_dist_local, int64_t n, int k, int64_t q_first,
                            int64_t q_count, int32_t *d_edges, double *d_weights, int64_t capacity, int64_t *n_edges, void *stream);
/* All phases on one GPU. Host outputs (any may be NULL): states[n][dim], edges[edge_capacity][2], valid[edge_capacity],
 * weights[edge_capacity] (edge length, the reference's double); n * k always suffices for edge_capacity. k <= 64. */
int b200mp_prm_build(b200mp_ctx *ctx, int robot, const b200mp_env *env, uint64_t seed, int64_t n, int k, int resolution, float *states,
                     int32_t *edges, uint8_t *valid, double *weights, int64_t edge_capacity, b200mp_roadmap_stats *stats);

/* ---- several GPUs of one node ------------------------------------------------------------------------------------- */
/* A b200mp_comm is one rank of a group of GPUs: a context plus an NCCL communicator (libnccl.so.2 is loaded when the first
 * group is made; without it these calls fail with B200MP_ERR_STATE, single-GPU use never needs it). Two ways to form a group:
 *   one process per GPU (MPI / torchrun style): rank 0 calls b200mp_comm_unique_id, the bytes reach the other processes by any
 *     means, every process calls b200mp_comm_create(ctx, id, rank, world);
 *   one process, one thread per GPU: b200mp_comm_create_all(ctxs, n, comms) (ncclCommInitAll), each thread then drives its rank.
 * b200mp_prm_build_multi is collective: every rank calls it with the same robot / seed / n / k / resolution and an environment
 * holding the same scene on its own context (b200mp_env_clone copies one across contexts). Rank r samples and validates a
 * contiguous slice of every round of the candidate stream (all-gather of the valid states), answers the nearest-neighbour
 * queries of a share of the states (whole leaves of the k-d index; ONE int32 max all-reduce of n k 4 bytes merges the tables),
 * lists and validates the edges of the states r B .. (r + 1) B - 1, B = ceil(n / world), and an all-gather leaves the whole
 * roadmap on every rank. The result is identical for every world size, and equal to b200mp_prm_build's. `neighbors` receives the
 * nearest-neighbour table the edges were derived from (ids ascending by (distance, id), self excluded). */
#define B200MP_UNIQUE_ID_BYTES 128
int b200mp_comm_unique_id(void *id, size_t bytes);
int b200mp_comm_create(b200mp_ctx *ctx, const void *unique_id, int rank, int world, b200mp_comm **out);
int b200mp_comm_create_all(b200mp_ctx *const *ctxs, int n, b200mp_comm **out /* [n] */);
void b200mp_comm_destroy(b200mp_comm *comm);
int b200mp_comm_rank(const b200mp_comm *comm, int *rank, int *world);
int b200mp_prm_build_multi(b200mp_comm *comm, int robot, const b200mp_env *env, uint64_t seed, int64_t n, int k, int resolution, int flags,
                           float *states, int32_t *edges, uint8_t *valid, double *weights, int32_t *neighbors /* [n][min(k, n - 1)], may be NULL */,
                           int64_t edge_capacity, b200mp_roadmap_stats *stats);
/* flags of b200mp_prm_build_multi. B200MP_PRM_INCREMENTAL: the graph stock PRM grows when the states are inserted in index order --
 * state i is connected to its k nearest among the states 0 .. i - 1 only (KStrategy queried before nn_->add, PRM.cpp:583,601), every
 * such pair is an edge (no symmetric union), checked neighbour -> new milestone like PRM::addMilestone (PRM.cpp:585-598). Slower than
 * the default (early states have few candidates, so their searches reach far), meant for graph-for-graph comparison with stock OMPL. */
enum { B200MP_PRM_INCREMENTAL = 1 };
/* A copy of `src` (primitive shapes, grid setting; scenes with point clouds or height fields are refused with
 * B200MP_ERR_CAPACITY) on another context, committed when `src` is. */
int b200mp_env_clone(const b200mp_env *src, b200mp_ctx *dst_ctx, b200mp_env **out);

/* FP32 FMA micro-benchmark used as the measured roofline denominator of the collision kernels:
 * runs `iters` dependent-chain FFMA blocks on every SM and returns achieved TFLOP/s. */
int b200mp_fp32_peak(b200mp_ctx *ctx, int iters, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* B200MP_H */
