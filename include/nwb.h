/* nwb.h - C ABI of the B200-native NAO whole-body validity / planning core.
 *
 * Drop-in boundary for the data-parallel hot path of burgetf/nao_wholebody_planning.  Every entry
 * point names the reference interface it replaces (file:line relative to the reference checkout):
 *   RP  = rrt_connect_planner/src/rrt_planner.cpp
 *   RPh = rrt_connect_planner/include/rrt_connect_planner/rrt_planner.h
 *   LP  = rrt_connect_planner/src/local_planner.cpp
 *   SCG = rrt_connect_planner/src/stable_config_generator.cpp
 *   NCK = rrt_connect_planner/src/nao_constraint_kinematics.cpp
 *   NPC = rrt_connect_planner/src/nao_planner_control.cpp
 *   AOC = rrt_connect_planner/src/articulated_object_constraints.cpp
 *   IKF = rrt_connect_planner/src/ikfast_TranslationDirection5D.cpp
 *
 * Conventions
 *   - A configuration is NWB_NJ = 21 doubles in the planner's joint order (RP:2003-2023), row-major.
 *   - Plain pointers and sizes only; the caller owns every buffer, the library owns nwb_ctx / nwb_tree.
 *   - Every function returns 0 (NWB_OK) or a negative NWB_E_* code; nothing exits or throws across
 *     the ABI (the reference calls exit(1) on I/O errors, SCG:205,218, RP:1440).
 *   - A context is bound to one CUDA device and one stream and is not re-entrant (the reference's
 *     planner objects are not either: RPh:137-138); use one context per GPU / per host thread.
 *   - `*_dev` variants take device pointers on the context's device and only enqueue work on the
 *     context's stream (no host synchronisation); call nwb_ctx_synchronize() to wait.
 *   - There is no CPU fallback: without a usable CUDA device nwb_ctx_create fails with NWB_E_CUDA.
 */
#ifndef NWB_H_
#define NWB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NWB_API __attribute__((visibility("default")))
#else
#define NWB_API
#endif

#define NWB_NJ 21          /* joints per configuration (RP:2001-2023)                       */
#define NWB_NFRAMES 29     /* link frames returned by nwb_fk_batch (models/nao_h25_rfoot)  */
#define NWB_ABI_VERSION 1

/* error codes */
#define NWB_OK 0
#define NWB_E_INVALID (-1)   /* bad argument (NULL, negative size, wrong length: NPC:242 has no check) */
#define NWB_E_CUDA (-2)      /* CUDA runtime error; text in nwb_last_error()                   */
#define NWB_E_NOMEM (-3)
#define NWB_E_IO (-4)
#define NWB_E_STATE (-5)     /* call sequence error (e.g. solve before start/goal are set)     */

/* enum Status { ADVANCED, REACHED, TRAPPED }  (RPh:46) - same numeric values */
#define NWB_ADVANCED 0
#define NWB_REACHED 1
#define NWB_TRAPPED 2

/* reason bits written by the validity operators */
#define NWB_REASON_COLLISION 1u   /* some enabled link pair / scene box in contact (RP:1257)   */
#define NWB_REASON_UNSTABLE 2u    /* projected COM not strictly inside the support hull (RP:1275) */

/* Reference quirks that change results.  A set bit reproduces the shipped
 * behaviour, a clear bit the evidently intended one. */
#define NWB_QUIRK_DB_KEEPS_INFEASIBLE 1u  /* C1  SCG:452-456 inverted test in the database branch   */
#define NWB_QUIRK_SIGNED_FOOT_ERROR 2u    /* C2  LP:283-297 signed max of the position error        */
#define NWB_QUIRK_DB_INDEX_FLOOR 4u       /* C3  RP:319-320 floor(U(0,n-1)): last row never drawn   */
#define NWB_QUIRKS_DEFAULT (NWB_QUIRK_SIGNED_FOOT_ERROR | NWB_QUIRK_DB_INDEX_FLOOR)

/* collision-pair set selector */
#define NWB_PAIRS_GROUP 0   /* isFeasible: group "whole_body_rfoot" filter, 108 pairs (RP:1250-1257) */
#define NWB_PAIRS_ALL 1     /* isPathValid: whole robot, 109 pairs (RP:1876)                        */

/* element type of device-resident configurations for the *_dev variants */
#define NWB_F64 0
#define NWB_F32 1

typedef struct nwb_ctx nwb_ctx;

/* Planner parameters: the private-namespace ROS parameters of NPC:20-51 /
 * planning_config/planning_parameters.yaml plus the constants hard-coded in the sources. */
typedef struct nwb_params {
  int32_t device;              /* CUDA device ordinal                                          */
  int32_t quirks;              /* NWB_QUIRK_* bit mask                                          */
  int32_t srdf_trim;           /* 1: trim stray blanks in SRDF link names (srdfdom), SRDF:230,236,301 */
  int32_t scale_origin;        /* support-polygon scaling about 0: sole origin, 1: the model's foot centre (default), 2: mean of the polygon vertices */
  double scale_sp;             /* planning_parameters.yaml:3 (0.8); RP:105-109                  */
  double ds_tolerance;         /* LP:293 (0.01 m)                                               */
  double ik_eps;               /* LP:182, NCK:127 (1e-3)                                        */
  int32_t ik_maxiter;          /* LP:182, NCK:127 (100)                                         */
  int32_t scg_max_samples;     /* planning_parameters.yaml:6 (5000)                             */
  int32_t scg_max_ik_iterations; /* planning_parameters.yaml:7 (5)                              */
  int32_t planner_max_iterations; /* planning_parameters.yaml:11 (8000)                         */
  double planner_step_factor;  /* planning_parameters.yaml:10 (0.1)                             */
  double pinv_eps;             /* KDL ChainIkSolverVel_pinv default (1e-5)                      */
  int32_t shortcut_waypoints;  /* RP:1396 (100)                                                 */
  int32_t shortcut_max_rounds; /* RP:1856 (500)                                                 */
  /* limb weights of the approach / manipulation plans (planning_parameters.yaml:13-21, NPC:27-32) */
  double r_arm_weight_approach, l_arm_weight_approach, legs_weight_approach;       /* 1.0, 1.0, 1.0 */
  double r_arm_weight_manipulate, l_arm_weight_manipulate, legs_weight_manipulate; /* 1.0, 0.0, 1.0 */
  int32_t manipulate_waypoints;   /* RP:1398 (4)   way-points per segment of an articulated path  */
  int32_t object_interp_points;   /* NPC:539,822 (20) intermediate points of the hand trajectory  */
  double manipulation_velocity;   /* RP:92 (0.02 m/s)                                             */
  /* generateTrajectory's time parameterisation (RP:1596-1618, MoveIt IterativeParabolicTimeParameterization):
   * the limits MoveIt reads from the robot model.  nao_description2 is not in the reference tree: the
   * velocity defaults are the NAO V4 URDF values quoted from memory (UNVERIFIED); the URDF carries no
   * acceleration limits, so MoveIt's default 1.0 rad/s^2 applies. */
  int32_t time_parameterization;  /* 1: apply it where the reference does (not for execute && articulated) */
  int32_t iptp_max_iterations;    /* 100  (IterativeParabolicTimeParameterization default)        */
  double iptp_max_time_change;    /* 0.01 (max_time_change_per_it)                                */
  double joint_max_velocity[NWB_NJ];
  double joint_max_acceleration[NWB_NJ];
} nwb_params;

/* Fill `p` with the reference defaults. */
NWB_API int nwb_params_default(nwb_params* p);

/* Library / ABI version (NWB_ABI_VERSION of the built library). */
NWB_API int nwb_abi_version(void);

/* Context = robot model + planning scene + parameters on one device.
 * Replaces the construction in NAO_PLANNER_CONTROL (NPC:87-104): PlanningScene from the robot
 * model, StableConfigGenerator / RRT_Planner with scale_sp.  The scene starts empty, as in NPC:90-92. */
NWB_API int nwb_ctx_create(const nwb_params* params, nwb_ctx** out);
NWB_API void nwb_ctx_destroy(nwb_ctx* ctx);
NWB_API const char* nwb_last_error(const nwb_ctx* ctx);   /* ctx may be NULL: last create error */
NWB_API int nwb_ctx_synchronize(nwb_ctx* ctx);
/* RRT_Planner::scale_support_polygon (RP:105-109) / StableConfigGenerator ctor (SCG:63). */
NWB_API int nwb_scale_support_polygon(nwb_ctx* ctx, double scale_sp);
/* Joint limits of the 21 planner joints (read from the RobotModel in RP:64-72). */
NWB_API int nwb_get_joint_limits(const nwb_ctx* ctx, double* min_out /*[21]*/, double* max_out /*[21]*/);
/* Raw CUDA stream handle (cudaStream_t) of the context, for callers that enqueue around it. */
NWB_API void* nwb_ctx_stream(nwb_ctx* ctx);
/* Make the context launch on a caller-owned stream (cudaStream_t; NULL restores the context's own).
 * Lets a host framework order these kernels with its own work and time them with its own events. */
NWB_API int nwb_ctx_set_stream(nwb_ctx* ctx, void* cuda_stream);

/* Planning scene: oriented boxes in the r_sole frame (moveit_msgs/CollisionObject BOX primitives as
 * the examples insert them: planning_example_simulation.cpp:310-357, initScene.cpp:120-158; all
 * commented out in the shipped node, so the default scene is empty like NPC:90-92).
 *   center[3], size[3] = full extents, rot[9] = row-major rotation (NULL = axis aligned). */
NWB_API int nwb_scene_clear(nwb_ctx* ctx);
NWB_API int nwb_scene_add_box(nwb_ctx* ctx, const double* center, const double* size, const double* rot);
/* n boxes at once, boxes15 [n][15] = {center[3], size[3], rot[9]}: one rebuild of the scene structures for the set.
 * Up to 16 boxes travel in the kernel parameters; larger scenes (up to 2^20 boxes) are kept in device memory under a
 * flattened bounding-volume hierarchy that the validity kernels stage through shared memory (exact capsule-box leaf
 * test either way: same verdicts as testing every box). */
NWB_API int nwb_scene_add_boxes(nwb_ctx* ctx, const double* boxes15, int64_t n);
NWB_API int nwb_scene_num_boxes(const nwb_ctx* ctx);

/* Batched state validity: collision-free AND statically stable, both tests always evaluated
 * (RRT_Planner::isConfigValid RP:701-745 -> isFeasible RP:1219-1287; identical copy SCG:77-159).
 *   q      [n][21]  configurations (host)
 *   valid  [n]      1 = feasible
 *   reason [n]      NWB_REASON_* bits, may be NULL
 *   margins[n][2]   {min proxy separation, signed distance COM -> support-hull boundary} in metres,
 *                   > 0 = free / stable; may be NULL (then the cheaper kernel variant runs)
 */
NWB_API int nwb_is_feasible_batch(nwb_ctx* ctx, const double* q, int64_t n, uint8_t* valid, uint8_t* reason,
                          float* margins);
/* Same for callers that keep their configurations in single precision: q [n][21] float (host).  The kernels
 * evaluate in fp32 either way (the fp64 input is rounded on load), so for float data the verdicts are those of
 * nwb_is_feasible_batch on the widened values; the host->device traffic, which bounds this entry point, halves. */
NWB_API int nwb_is_feasible_batch_f32(nwb_ctx* ctx, const float* q, int64_t n, uint8_t* valid, uint8_t* reason,
                                      float* margins);
/* Same, device-resident.  q_dev is [n][21] of `dtype` (NWB_F64 | NWB_F32). */
NWB_API int nwb_is_feasible_batch_dev(nwb_ctx* ctx, const void* q_dev, int dtype, int64_t n,
                              uint8_t* valid_dev, uint8_t* reason_dev, float* margins_dev);

/* Collision-only validity over the whole robot, the predicate PlanningScene::isPathValid applies to
 * every way-point (RP:1876,1905; the stability predicate is never registered: RP:233). */
NWB_API int nwb_is_state_colliding_batch(nwb_ctx* ctx, const double* q, int64_t n, uint8_t* colliding,
                                 float* min_separation /* [n], may be NULL */);

/* Forward kinematics of the whole body rooted at r_sole + centre of mass
 * (MoveIt RobotState FK, called RP:715-720; hrl_kinematics computeCOM, called RP:1275).
 *   link_poses [n][NWB_NFRAMES][12]  rows of [R | p] (3x4, row-major) in the r_sole frame, may be NULL
 *   com        [n][3]                may be NULL
 * Computed in fp32 on the device and widened; agrees with the fp64 oracle to 1e-5 m. */
NWB_API int nwb_fk_batch(nwb_ctx* ctx, const double* q, int64_t n, double* link_poses, double* com);
/* Name of link frame `i` (0 <= i < NWB_NFRAMES), e.g. "l_sole"; NULL if out of range. */
NWB_API const char* nwb_frame_name(int i);

/* ---- tree + nearest neighbour ------------------------------------------------------------------
 * Device-resident replacement of struct Node / struct Tree (RPh:30-46).  A node is its configuration
 * plus predecessor_index; its index is its insertion position (RRT_Planner::addConfigtoTree, RP:749-783).
 * The library owns the nwb_tree; it belongs to the context it was created on. */
typedef struct nwb_tree nwb_tree;
NWB_API int nwb_tree_create(nwb_ctx* ctx, int64_t capacity_hint, nwb_tree** out);
NWB_API void nwb_tree_destroy(nwb_tree* tree);
NWB_API int nwb_tree_clear(nwb_tree* tree);                      /* RP:1110-1114: nodes.clear()            */
NWB_API int64_t nwb_tree_size(const nwb_tree* tree);             /* Tree::num_nodes                        */
/* Append n nodes (q [n][21], predecessor [n] or NULL = -1); indices continue from nwb_tree_size(). */
NWB_API int nwb_tree_add(nwb_tree* tree, const double* q, const int32_t* predecessor, int64_t n);
NWB_API int nwb_tree_add_dev(nwb_tree* tree, const double* q_dev, const int32_t* predecessor_dev, int64_t n);
/* Read back nodes [first, first+n): configurations and/or predecessors (either may be NULL). */
NWB_API int nwb_tree_get(nwb_tree* tree, int64_t first, int64_t n, double* q_out, int32_t* predecessor_out);
/* RRT_Planner::findNearestNeighbour (RP:336-419) for nq query configurations against one tree:
 * idx[k] = argmin_i sqrt(sum_j (q_k[j] - node_i[j])^2), first strict minimum (lowest index on ties),
 * -1 if no node is closer than 10000.0 (RP:347); dist[k] = that distance (may be NULL).
 * Exact: candidates are re-evaluated in double with the reference's rounding sequence. */
NWB_API int nwb_nearest_neighbour_batch(nwb_tree* tree, const double* q, int64_t nq, int32_t* idx, double* dist);
NWB_API int nwb_nearest_neighbour_batch_dev(nwb_tree* tree, const double* q_dev, int64_t nq, int32_t* idx_dev,
                                            double* dist_dev);

/* Diagnostic for the tests: an id of the scan thread that evaluates `node` when nq queries are passed (nodes with
 * equal ids are evaluated one after the other by the same thread, in index order).  No reference counterpart. */
NWB_API int64_t nwb_nn_scan_owner(const nwb_tree* tree, int64_t nq, int64_t node);
/* Diagnostic for the tests: the raw tensor-core accumulator of the batched scan for the first 128 nodes,
 * tile_out [128 queries][128 nodes] = |n|^2 - 2 q.n evaluated on bf16 hi/mid splits (rows >= nq are padding). */
NWB_API int nwb_nn_debug_tensor_tile(nwb_tree* tree, const double* q, int64_t nq, float* tile_out);

/* ---- local planner operators ---------------------------------------------------------------------
 * weights: the 21 joint weights of RRT_Planner::setJointWeights (RP:307-311), NULL = all 1.0 (RP:76-78).
 * step:    max_step_size of solveQuery (planner_step_factor, 0.1). */

/* generate_q_new (RP:424-494) followed by enforce_DS_Constraints (RP:498-593), combined the way
 * extendTree does (RP:807-818): status[i] = NWB_TRAPPED if the projection failed (joint limits or
 * left-leg IK), else NWB_ADVANCED if the left leg was re-solved, else generate_q_new's status.
 * q_out[i] is written only when status[i] != NWB_TRAPPED.  ds_status (may be NULL) = the
 * enforce_DS_Constraints result on its own. */
NWB_API int nwb_steer_project_batch(nwb_ctx* ctx, const double* q_near, const double* q_target, int64_t n,
                                    const double* weights, double step, double* q_out, int32_t* status,
                                    int32_t* ds_status);

/* The connectTree walk (RP:926-1031) for n_edges independent edges: from q_start[e] step towards
 * q_connect[e] (steer -> project -> isConfigValid -> append, each step starting from the projected
 * previous node, RP:1004) until NWB_REACHED or NWB_TRAPPED.  nodes_out [n_edges][max_steps][21]
 * receives the appended nodes, n_added[e] their count (rows past n_added[e] are not written: only the appended rows
 * are copied back); status[e] = -1 if max_steps ran out first. */
NWB_API int nwb_connect_batch(nwb_ctx* ctx, const double* q_start, const double* q_connect, int64_t n_edges,
                              const double* weights, double step, int32_t max_steps, double* nodes_out,
                              int32_t* n_added, int32_t* status);

/* interpolateWaypoints (RP:1677-1774: k intermediate + 2 end states) + PlanningScene::isPathValid
 * (RP:1876,1905: every way-point collision-free over the whole robot, no stability predicate) for
 * n_edges candidate edges qa[e] -> qb[e].  ok[e] = 1 if valid; first_bad[e] (may be NULL) = index of
 * the first colliding way-point or -1. */
NWB_API int nwb_is_path_valid_batch(nwb_ctx* ctx, const double* qa, const double* qb, int64_t n_edges, int32_t k,
                                    uint8_t* ok, int32_t* first_bad);

/* StableConfigGenerator::sample_stable_configs, database branch (SCG:278-480), for the sample indices
 * [first, first + count): uniform sample within the joint limits (LHipYawPitch = 0), hip FK, left-leg
 * IK (NCK:71-197), validity; kept rows (quirk C1 selects which) are returned in ascending sample
 * index.  The random stream is counter based (Philox4x32-10 keyed by `seed`, counter = sample index),
 * so the accepted set does not depend on how the index range is split over calls or GPUs.
 *   rows_out [cap][21], index_out [cap] (may be NULL), *n_out = number accepted (may exceed cap: then
 *   only the first cap rows in index order ... are returned and the call reports NWB_E_NOMEM),
 *   stage_out [count] (may be NULL): 0 = IK failed, 1 = infeasible, 2 = feasible. */
NWB_API int nwb_sample_stable_configs(nwb_ctx* ctx, uint64_t seed, uint64_t first, int64_t count, int32_t max_ik_iter,
                                      double* rows_out, int64_t cap, int64_t* n_out, uint64_t* index_out,
                                      uint8_t* stage_out);
/* Device-resident form for throughput runs: rows in arrival order (not sorted), *n_out_dev is an
 * unsigned 64-bit counter the caller zeroes.  rows_dev / index_dev / n_out_dev may be PEER memory (another
 * rank's nwb_dev_alloc block opened with nwb_ipc_import): the counter is bumped with a system-scope atomic,
 * so all ranks of a sharded generation can append into one rank's database - no separate gather. */
NWB_API int nwb_sample_stable_configs_dev(nwb_ctx* ctx, uint64_t seed, uint64_t first, int64_t count,
                                          int32_t max_ik_iter, double* rows_dev, uint64_t* index_dev, int64_t cap,
                                          unsigned long long* n_out_dev, uint8_t* stage_dev);
NWB_API int nwb_is_path_valid_batch_dev(nwb_ctx* ctx, const double* qa_dev, const double* qb_dev, int64_t n_edges,
                                        int32_t k, uint8_t* ok_dev, int32_t* first_bad_dev);
NWB_API int nwb_steer_project_batch_dev(nwb_ctx* ctx, const double* q_near_dev, const double* q_target_dev, int64_t n,
                                        const double* weights, double step, double* q_out_dev, int32_t* status_dev,
                                        int32_t* ds_status_dev);

/* ---- RRT-Connect driver --------------------------------------------------------------------------
 * Result of one planning query (the response fields of Compute_Motion_Plan.srv:16-29 that come out
 * of RRT_Planner::solveQuery, plus what writePath needs to rebuild the path). */
typedef struct nwb_plan_result {
  int32_t start_valid;          /* RRT_Planner::setStartGoalConfigs, RP:241-246                    */
  int32_t goal_valid;           /* RP:260-265                                                       */
  int32_t success;              /* a path was found (REACHED), RP:1087 / RP:1147                    */
  int32_t iterations;           /* solveQuery loop iterations used (RP:1067)                        */
  int32_t num_nodes_generated;  /* T_start.num_nodes + T_goal.num_nodes (RP:1108)                   */
  int32_t nodes_start;
  int32_t nodes_goal;
  int32_t connector;            /* 1: goal tree connected to start tree, 2: the reverse (RP:1101,1161) */
  int32_t bridge_other;         /* q_near.index in the connecting tree (RP:1320,1325)               */
  int32_t raw_path_len;         /* way-points of the raw solution path (RP:1383-1391)               */
} nwb_plan_result;

/* RRT_Planner::setStartGoalConfigs + solveQuery (RP:227-302, 1039-1212) + the path extraction of
 * writePath (RP:1303-1392) for nq independent queries advanced in lock step on the device.
 *   starts, goals [nq][21]; db [n_db][21] = the stable-configuration database (load_DS_database,
 *   RP:121-191); weights [21] or NULL; seeds [nq]: query k draws its database row of iteration i
 *   from Philox4x32-10(seed_k; i) (RRT_Planner::getRandomStableConfig RP:316-332 with an injected
 *   stream; quirk C3 selects floor(U*(n-1)) or floor(U*n)); max_iter, step: solveQuery arguments
 *   (8000, 0.1); tree_capacity: nodes per tree (0 = 4096).
 *   results [nq]; raw_paths [nq][raw_path_capacity][21] (may be NULL) receives each raw path. */
NWB_API int nwb_solve_query_batch(nwb_ctx* ctx, const double* starts, const double* goals, int64_t nq, const double* db,
                                  int32_t n_db, const double* weights, const uint64_t* seeds, int32_t max_iter, double step,
                                  int32_t tree_capacity, nwb_plan_result* results, double* raw_paths,
                                  int32_t raw_path_capacity);
/* RRT_Planner::path_shortcutter, way-point selection (RP:1830-1951): returns the number of way-points
 * written to out (<= cap), -100 if the 500-round budget ran out, or an NWB_E_* code.  Every round's
 * candidate edges are checked in one batch of nwb_is_path_valid_batch. */
NWB_API int nwb_shortcut_path(nwb_ctx* ctx, const double* raw, int32_t n_raw, double* out, int32_t cap,
                              int32_t* edges_checked);
/* RRT_Planner::interpolateShortPathWaypoints (RP:1956-1994): k points per segment; returns the row
 * count (n_path-1)*(k+1)+1 (rows beyond cap are not written). */
NWB_API int nwb_interpolate_path(const double* path, int32_t n_path, int32_t k, double* out, int32_t cap);

/* ---- multi-GPU result gather without a collective --------------------------------------------------
 * Independent validity batches shard over the GPUs of one box with no data-path exchange; only the
 * 1 B/config verdicts have to reach every rank.  Instead of a separate all-gather, the validity
 * kernel itself stores each verdict into the gathered array of every rank through peer-mapped
 * memory (NVLink 5 / NVSwitch).  One process per GPU: every rank allocates its gathered array with
 * nwb_dev_alloc, exports it (nwb_ipc_export, 64-byte handle, exchanged by the host framework),
 * imports the others (nwb_ipc_import) and registers the n_ranks base pointers (own one included, in
 * rank order) with nwb_set_gather_peers; from then on nwb_is_feasible_batch_dev also writes
 * verdict i to gathered[r][my_offset + i] on every rank r.  A rank may read its gathered array after
 * its own stream and a cross-rank barrier have completed.  n_ranks = 0 switches the fused gather off. */
NWB_API int nwb_dev_alloc(nwb_ctx* ctx, size_t bytes, void** out);
NWB_API int nwb_dev_free(nwb_ctx* ctx, void* p);
NWB_API int nwb_ipc_export(nwb_ctx* ctx, void* dev_ptr, unsigned char* handle64);
NWB_API int nwb_ipc_import(nwb_ctx* ctx, const unsigned char* handle64, void** out);
NWB_API int nwb_ipc_close(nwb_ctx* ctx, void* p);
NWB_API int nwb_set_gather_peers(nwb_ctx* ctx, void* const* gathered, int32_t n_ranks, int64_t my_offset);

/* ---- on-disk formats + the database service ----------------------------------------------------------
 * The reference's persistent artefacts are text files: one configuration per line, values printed by
 * `ofstream << double << " "` (6 significant digits, trailing blank; SCG:463-468, RP:1444-1452) and
 * parsed with strtod (RRT_Planner::load_DS_database, RP:121-191). */
NWB_API int nwb_write_configs(const char* path, const double* rows, int64_t n_rows);
/* Reads up to cap rows of 21 values; *n_rows = rows in the file.  NWB_E_IO if it cannot be opened
 * (the reference prints a message and carries on, RP:140,189). */
NWB_API int nwb_load_ds_database(const char* path, double* rows, int64_t cap, int64_t* n_rows);

/* rrt_planner_msgs/srv/Generate_DS_Configs.srv:1-5 */
typedef struct nwb_generate_ds_configs_request {
  int32_t max_samples;
  int32_t max_ik_iterations;
} nwb_generate_ds_configs_request;
typedef struct nwb_generate_ds_configs_response {
  int32_t num_configs_generated;
} nwb_generate_ds_configs_response;
/* NAO_PLANNER_CONTROL::generate_ds_database (NPC:137-147): sample max_samples configurations, write
 * the accepted ones to `path` (the node's ssc_double_support file), report their number.  `seed`
 * replaces the reference's clock-seeded generator.  Like the service callback it reports failures
 * through the return code instead of exiting (SCG:205,218). */
NWB_API int nwb_generate_ds_database(nwb_ctx* ctx, const nwb_generate_ds_configs_request* req,
                                     nwb_generate_ds_configs_response* resp, const char* path, uint64_t seed);

/* ---- right arm, articulated objects, goal configurations ----------------------------
 * Hand poses are 6 doubles: position of the grasp point and direction of the hand z-axis, in the
 * right-sole frame unless stated otherwise (Compute_Goal_Config.srv:1-4). */

/* The generated right-arm solver IKSolver::ik (IKF:341-1517, entry IKF:1521) for n targets given in ITS
 * torso frame: targets [n][6] = position + unit direction; solutions [n][8][5] (RShoulderPitch,
 * RShoulderRoll, RElbowYaw, RElbowRoll, RWristYaw, each in (-pi, pi]), counts [n].  Solved in closed
 * form on the device (csrc/arm_core.cuh); the solution SET equals the reference solver's. */
NWB_API int nwb_arm_ik_batch(nwb_ctx* ctx, const double* targets, int64_t n, double* solutions, int32_t* counts);
/* ConstraintKinematics::computeRightArmIK (NCK:224-506 with getHipPose NCK:512-544) for n whole-body rows
 * q [n][21]: hip pose from the right leg, then
 *   direction != 0 (NCK:374-506): hand pose -> the solver's torso frame, all solutions, strict joint limits, the one
 *     with the largest compute_ergonomics_index (NCK:612-659);
 *   direction == (0,0,0) exactly (NCK:291-372): KDL NR(100, 1e-3) on the right-arm chain towards
 *     Frame(position in the hip frame) from up to 5 random seeds inside the arm limits; row i draws its seeds
 *     from Philox(seed; counter = first_index + i, stream 3) in place of the class's clock-seeded rng_;
 *     ergonomics stays 0.
 * ok [n], arm [n][5], ergonomics [n]. */
NWB_API int nwb_goal_arm_ik_batch(nwb_ctx* ctx, const double* q, int64_t n, const double* hand_pose6, uint64_t seed,
                                  uint64_t first_index, uint8_t* ok, double* arm, double* ergonomics);
/* ArticulatedObjectConstraints::computeDrawerTrajectoryPoints (kind 0, AOC:105-122) /
 * computeDoorTrajectoryPoints (kind 1, AOC:126-189; door4 = {rotation_radius, rotation_angle [deg],
 * relative_angle [deg], clearance_handle} as NPC:816-820): out [num_interp_points + 2][6]. */
NWB_API int nwb_hand_trajectory(int32_t kind, const double* start6, const double* goal6, int32_t num_interp_points,
                                const double* door4, double* out);
/* RRT_Planner::enforce_MA_Constraints (RP:597-696) for n independent (q_near, q_new_ds) pairs against
 * the trajectory points pts [n_points][6]: start_tree selects the index rule of "start_tree" /
 * "goal_tree"; near_hand_index [n] = q_near.cart_hand_pose_index.  status[i]: NWB_REACHED (hand already
 * on the desired point, AOC:204-262), NWB_ADVANCED (right arm re-solved, AOC:313-449) or NWB_TRAPPED;
 * q_out[i] written unless TRAPPED. */
NWB_API int nwb_enforce_ma_batch(nwb_ctx* ctx, const double* pts, int32_t n_points, int32_t start_tree,
                                 const int32_t* near_hand_index, const double* q_near, const double* q_new_ds, int64_t n,
                                 double* q_out, int32_t* status);
/* Goal branch of StableConfigGenerator::sample_stable_configs (SCG:278-436), one outcome per sample
 * index in [first, first + count): stage [count] (0 rejected, 1 both IKs solved but infeasible,
 * 2 accepted), ergonomics [count] and rows [count][21] (either may be NULL). */
NWB_API int nwb_goal_sample_batch(nwb_ctx* ctx, uint64_t seed, uint64_t first, int64_t count, const double* hand_pose6,
                                  int32_t max_ik_iter, uint8_t* stage, double* ergonomics, double* rows);
/* sample_stable_configs(max_samples, max_ik_iter, t, true, file) + sort_goal_configs_by_ergonomy
 * (SCG:176-507, 510-661): the first six accepted samples in sample order, passed through the goal
 * file's text format (6 significant digits) and sorted by descending ergonomics index; row 0 is what
 * loadGoalConfig (SCG:666-756) returns.  rows6 [6][21], ergonomics6 [6], index6 [6] (may be NULL),
 * *n_found <= 6, *samples_drawn (may be NULL) = loop iterations the serial reference would have used. */
NWB_API int nwb_sample_goal_configs(nwb_ctx* ctx, uint64_t seed, const double* hand_pose6, int32_t max_samples,
                                    int32_t max_ik_iter, double* rows6, double* ergonomics6, uint64_t* index6,
                                    int32_t* n_found, int64_t* samples_drawn);
/* nwb_solve_query_batch with RRT_Planner::activate_drawer_constraint / activate_door_constraint in force
 * (RP:195-223): hand_trajectories [nq][n_points][6], one per query; nodes carry cart_hand_pose_index
 * (RPh:35), extendTree / connectTree apply enforce_MA_Constraints (RP:829-837, 917-973). */
NWB_API int nwb_solve_query_constrained_batch(nwb_ctx* ctx, const double* starts, const double* goals, int64_t nq,
                                              const double* db, int32_t n_db, const double* weights, const uint64_t* seeds,
                                              int32_t max_iter, double step, int32_t tree_capacity,
                                              const double* hand_trajectories, int32_t n_points, nwb_plan_result* results,
                                              double* raw_paths, int32_t raw_path_capacity);

/* ---- planning services (NPC:150-870, the .srv files of rrt_planner_msgs) ------------------------------------------
 * Request / response structs mirror the .srv fields one to one; moveit_msgs/DisplayTrajectory is
 * flattened into caller-provided arrays.  Like the callbacks, the functions report planning failure
 * through the response flags and return NWB_OK; a negative return is an argument / device error.
 * The stable-configuration database is the context's (nwb_ctx_set_ds_database, or the file loaded /
 * generated through nwb_ctx_load_ds_database / nwb_generate_ds_database, NPC:98,144).  `seed` replaces
 * the clock-seeded generators: goal sampling uses seed (approach) and seed + 1 (manipulation), the two
 * solveQuery calls seed + 2 and seed + 3. */
typedef struct nwb_point3 {  /* geometry_msgs/Point, geometry_msgs/Vector3 */
  double x, y, z;
} nwb_point3;
typedef struct nwb_trajectory {   /* moveit_msgs/DisplayTrajectory -> joint_trajectory.points */
  int32_t capacity;               /* in:  rows the two arrays can hold                                   */
  int32_t n_points;               /* out: way-points of the plan (rows beyond capacity are not written)  */
  double* positions;              /* [capacity][21]                                                      */
  double* time_from_start;        /* [capacity] seconds: nwb_time_parameterize of the way-points, or - for an
                                   * executed articulated plan / time_parameterization = 0 - the stamps
                                   * 2.0 + i * increment of RP:1532-1553                                   */
} nwb_trajectory;

NWB_API int nwb_ctx_set_ds_database(nwb_ctx* ctx, const double* rows, int64_t n_rows);
NWB_API int nwb_ctx_load_ds_database(nwb_ctx* ctx, const char* path);   /* RRT_Planner::load_DS_database, RP:121-191 */
NWB_API int64_t nwb_ctx_ds_database_size(const nwb_ctx* ctx);

/* trajectory_processing::IterativeParabolicTimeParameterization::computeTimeStamps as called by
 * RRT_Planner::generateTrajectory (RP:1610-1611) on n_points way-points [n_points][21]: velocity
 * constraints, then the iterative forward / backward expansion of the intervals until every joint's
 * parabolic-blend acceleration is within its limit.  time_from_start [n_points] receives the cumulative
 * durations + 2.0 s (RP:1614-1618).  Host-side and serial, as in the reference.  MoveIt is not vendored:
 * restated from its published algorithm - PARITY UNPINNED (no shipped file carries time stamps). */
NWB_API int nwb_time_parameterize(const nwb_ctx* ctx, const double* positions, int32_t n_points, double* time_from_start);

typedef struct nwb_compute_goal_config_request {   /* Compute_Goal_Config.srv:1-4 */
  nwb_point3 right_hand_position;
  nwb_point3 right_hand_direction;
} nwb_compute_goal_config_request;
typedef struct nwb_compute_goal_config_response {  /* Compute_Goal_Config.srv:6-8 */
  double wb_goal_config[NWB_NJ];
  int32_t num_goal_configs_found;                  /* not in the .srv: the callback drops it (NPC:170) */
} nwb_compute_goal_config_response;
/* NAO_PLANNER_CONTROL::compute_goal_config (NPC:150-181) */
NWB_API int nwb_compute_goal_config(nwb_ctx* ctx, const nwb_compute_goal_config_request* req,
                                    nwb_compute_goal_config_response* resp, uint64_t seed);

typedef struct nwb_compute_motion_plan_request {   /* Compute_Motion_Plan.srv:1-12 */
  nwb_point3 right_hand_position;
  nwb_point3 right_hand_direction;
  const double* wb_start_config;                   /* float64[]: must hold 21 values (NPC:242 does not check) */
  int32_t wb_start_config_len;
  uint8_t execute_trajectory;
} nwb_compute_motion_plan_request;
typedef struct nwb_compute_motion_plan_response {  /* Compute_Motion_Plan.srv:14-29 */
  double wb_goal_config[NWB_NJ];
  double time_goal_config;
  nwb_trajectory motion_plan;
  uint32_t num_nodes_generated;
  double search_time;
  uint8_t success;
} nwb_compute_motion_plan_response;
/* NAO_PLANNER_CONTROL::compute_motion_plan (NPC:183-336) */
NWB_API int nwb_compute_motion_plan(nwb_ctx* ctx, const nwb_compute_motion_plan_request* req,
                                    nwb_compute_motion_plan_response* resp, uint64_t seed);

/* n independent compute_motion_plan requests (the grasp-pose evaluation loop of
 * planning_example_simulation.cpp:770-900 calls the service once per pose): all goal samplings share
 * three launches, all RRT-Connect queries run in one, all shortcut rounds in shared edge-check batches;
 * the per-request host finishing (goal-file text round trip, time parameterisation) runs on host threads.
 * Request k uses seeds[k] exactly as nwb_compute_motion_plan would, so resps[k] equals the single call's
 * response except for the two wall-clock fields (batch times). */
NWB_API int nwb_compute_motion_plan_batch(nwb_ctx* ctx, const nwb_compute_motion_plan_request* reqs,
                                          nwb_compute_motion_plan_response* resps, int64_t n, const uint64_t* seeds);

typedef struct nwb_manipulation_plan_response {    /* Compute_Linear_Manipulation_Plan.srv:20-46 == Circular:27-53 */
  nwb_trajectory motion_plan_approach;
  nwb_trajectory motion_plan_manipulation;
  double wb_goal_config_approach[NWB_NJ];
  double time_goal_config_approach;
  double wb_goal_config_manipulate[NWB_NJ];
  double time_goal_config_manipulate;
  uint32_t num_nodes_generated_approach;
  uint32_t num_nodes_generated_manipulate;
  double search_time_approach;
  double search_time_manipulate;
  uint8_t success_approach;
  uint8_t success_manipulation;
} nwb_manipulation_plan_response;

typedef struct nwb_compute_linear_manipulation_plan_request {   /* Compute_Linear_Manipulation_Plan.srv:1-18 */
  nwb_point3 right_hand_position;
  nwb_point3 right_hand_direction;
  const double* wb_start_config;
  int32_t wb_start_config_len;
  double relative_angle;               /* degrees */
  double object_translation_length;    /* metres  */
  uint8_t execute_trajectory;
} nwb_compute_linear_manipulation_plan_request;
/* NAO_PLANNER_CONTROL::compute_linear_manipulation_plan (NPC:338-588): approach plan, then the drawer
 * pull under activate_drawer_constraint (RP:195-206). */
NWB_API int nwb_compute_linear_manipulation_plan(nwb_ctx* ctx, const nwb_compute_linear_manipulation_plan_request* req,
                                                 nwb_manipulation_plan_response* resp, uint64_t seed);

typedef struct nwb_compute_circular_manipulation_plan_request { /* Compute_Circular_Manipulation_Plan.srv:1-25 */
  nwb_point3 right_hand_position;
  nwb_point3 right_hand_direction;
  const double* wb_start_config;
  int32_t wb_start_config_len;
  double relative_angle;               /* degrees */
  double rotation_radius;
  double rotation_angle;               /* degrees */
  double clearance_handle;
  uint8_t execute_trajectory;
} nwb_compute_circular_manipulation_plan_request;
/* NAO_PLANNER_CONTROL::compute_circular_manipulation_plan (NPC:590-870): approach plan, then the door
 * swing under activate_door_constraint (RP:209-223). */
NWB_API int nwb_compute_circular_manipulation_plan(nwb_ctx* ctx, const nwb_compute_circular_manipulation_plan_request* req,
                                                   nwb_manipulation_plan_response* resp, uint64_t seed);

/* Pinned host memory helpers: buffers from here make the host<->device copies of the batch
 * operators run at full PCIe speed. */
NWB_API void* nwb_host_alloc(size_t bytes);
NWB_API void nwb_host_free(void* p);

/* FP32-pipe micro-benchmark: dependent FFMA chains on every SM; returns achieved TFLOP/s.
 * Used as the measured roofline denominator for the FP32-bound kernels. */
NWB_API int nwb_ffma_peak(nwb_ctx* ctx, int iters, double* tflops_out, double* ms_out);

/* Device time of the most recent batch operator on this context (CUDA events on the context's
 * stream around the kernel launches only), in milliseconds; and the number of kernels launched. */
NWB_API int nwb_last_kernel_ms(nwb_ctx* ctx, double* ms_out, int64_t* launches_out);

/* Switches the timing events of nwb_last_kernel_ms off (0) or on (non-zero; the default).  A caller that pipelines
 * operator calls back to back on one stream saves the stream time of two event records per call (measured: 5 us per
 * 131 us validity launch); nwb_last_kernel_ms then reports -1 ms and the launch count only. */
NWB_API int nwb_set_kernel_timing(nwb_ctx* ctx, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* NWB_H_ */
