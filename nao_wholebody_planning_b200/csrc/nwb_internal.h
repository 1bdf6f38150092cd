// nwb_internal.h - host-side context and kernel launcher prototypes (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/nwb.h"
#include "nao_model_gen.h"

namespace nwb {

constexpr int kMaxPairs = 112;
constexpr int kMaxRuns = 48;
constexpr int kMaxBoxes = 16;              // boxes that travel inline in the kernel parameters
constexpr int kCulledMinBoxes = 2;        // below this the inline boxes of the kernel parameters are faster
constexpr int kStagedNodes = 1024;         // hierarchy nodes a CTA of the batched validity kernel stages in shared memory
constexpr int kSceneBoxLimit = 1 << 20;    // boxes of a scene held in global memory under the hierarchy

// One collision pair inside a run: B's point-table slot and constants.
struct PairEntry {
  int off_b;     // float slot of B's first point
  float e;       // |b_B - a_B|^2
  float inv_e;   // 1/e (0 for spheres)
  float rsum;    // r_A + r_B
};
// Pairs sharing the same first proxy A, so A is loaded from the point table once.
struct PairRun {
  int off_a;
  float a;       // |b_A - a_A|^2
  float inv_a;
  int first;     // index of the first PairEntry
  int count;
};
struct alignas(16) SceneBox {   // oriented box in the r_sole frame (64 B: four 16-byte loads from global / shared memory)
  float c[3];
  float h[3];      // half extents
  float r[9];      // row-major rotation, columns = box axes
  float pad;
};
// Scenes of more than kMaxBoxes boxes live in global memory under a flattened bounding-volume hierarchy: nodes in
// depth-first order, `skip` = index of the next node when this subtree is culled (stack-free walk), a leaf holds ONE box.
struct alignas(16) SceneNode {
  float lo[3];
  int skip;
  float hi[3];
  int box;         // >= 0: leaf, index into the box array; -1: inner node
};
struct SceneView {
  const SceneNode* nodes;   // device memory, nullptr when the scene is inline (<= kMaxBoxes boxes)
  const SceneBox* boxes;
  int n_nodes;
  int n_boxes;
};
struct EnvLink {    // a robot proxy tested against the scene boxes
  int off;
  float inv_a;     // 1/|b-a|^2, 0 for spheres
  float radius;
  int is_sphere;
  float half_len;  // |b-a|/2: with the radius, the bounding sphere of the capsule about its mid point
};

// Passed to the kernels by value (__grid_constant__): lives in the constant bank, per launch,
// so contexts with different parameters never share global state.
struct ValidityTables {
  float foot_right[nwbm::kFootN][2];   // scaled right-foot polygon, r_sole frame
  float foot_left[nwbm::kFootN][2];    // scaled left-foot polygon, l_sole frame
  int n_runs_cc, n_runs_cs, n_runs_ss; // runs are stored in this order
  int n_boxes, n_env_links;
  int env_cull;                        // 1: skip the exact capsule-box test where a warp agrees the bounding sphere is clear
  PairRun runs[kMaxRuns];
  PairEntry entries[kMaxPairs];
  SceneBox boxes[kMaxBoxes];            // inline scene (n_boxes <= kMaxBoxes), else n_boxes = 0 and `scene` holds it
  EnvLink env_links[nwbm::kNumGeom];
  SceneView scene;                      // large scene: BVH + boxes in global memory (owned by the context)
  SceneView culled;                     // the same hierarchy for every scene of >= kCulledMinBoxes boxes: read by the
                                        // verdict-only batch kernel, which culls and compacts the exact tests
};

// Fused result gather: besides its own verdict array, the validity kernel stores each verdict into
// the gathered array of every rank (peer memory over NVLink / NVSwitch, mapped through CUDA IPC).
struct GatherPeers {
  uint8_t* ptr[8];     // base of each rank's gathered array (own rank included)
  int n;               // 0 = no fused gather
  int64_t offset;      // this rank's first element inside the gathered arrays
};

struct Ctx {
  nwb_params params;
  GatherPeers gather{};
  int device = 0;
  cudaStream_t stream = nullptr;      // stream the operators launch on (own_stream unless overridden)
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timing = true;                 // record ev0 / ev1 around the kernels of every operator (nwb_set_kernel_timing)
  cudaStream_t copy_stream = nullptr;                     // second stream: H2D copies overlap the kernels
  cudaEvent_t in_done[2] = {nullptr, nullptr}, k_done[2] = {nullptr, nullptr};
  std::string err;
  std::vector<SceneBox> boxes;
  std::vector<SceneNode> bvh;                          // host copy of the hierarchy (scenes of more than kMaxBoxes boxes)
  void* d_scene = nullptr; size_t d_scene_bytes = 0;   // [nodes | boxes] on the device
  bool scene_dirty = false;                            // boxes changed since the last upload (bulk adds upload once)
  ValidityTables tab_group;   // isFeasible pair set
  ValidityTables tab_all;     // isPathValid pair set
  // scratch device buffers for the host-pointer entry points (grown on demand)
  void* d_in = nullptr; size_t d_in_bytes = 0;
  void* d_out = nullptr; size_t d_out_bytes = 0;
  void* h_stage = nullptr; size_t h_stage_bytes = 0;   // pinned staging for pageable callers
  void* d_plan = nullptr; size_t d_plan_bytes = 0;     // scratch of the planner operators
  void* d_forest = nullptr; size_t d_forest_bytes = 0; // pooled forest of the RRT driver (grow-only)
  void* d_rrt = nullptr; size_t d_rrt_bytes = 0;       // control state / staging of the RRT driver (grow-only)
  int32_t* h_flag = nullptr;                           // pinned word the RRT driver polls
  std::vector<double> ds_db;                           // stable-configuration database of the services (RP:121-191)
  std::vector<double> h_raw;                           // raw-path staging of the services (kept across calls)
  uint64_t ds_db_version = 0;                          // bumped whenever ds_db changes
  double* d_dbcache = nullptr; size_t d_dbcache_rows = 0; uint64_t d_dbcache_version = ~0ull;   // device copy of ds_db
  void* d_work = nullptr; size_t d_work_bytes = 0;     // work list of the two-phase sampler (grow-only)
  void* h_pin = nullptr; size_t h_pin_bytes = 0;       // pinned staging of the fused RRT call (inputs out, results back)
  double last_ms = 0;
  int64_t last_launches = 0;
  int64_t last_nn_passes = 0;        // passes over the node array of the last nearest-neighbour call
  int sm_count = 0;
  bool force_table_pairs = false;   // NWB_FORCE_TABLE_PAIRS=1: use the table-driven pair loops (A/B measurements)
};

// Timing events around an operator's kernels: an event between two back-to-back launches costs ~2.5 us of stream time
// (measured: 131.1 -> 136.3 us per headline launch with the pair), so callers that pipeline launches switch them off.
inline cudaError_t ev_record(Ctx* ctx, cudaEvent_t ev, cudaStream_t s) { return ctx->timing ? cudaEventRecord(ev, s) : cudaSuccess; }

void build_tables(Ctx* ctx);

// kernel launchers (validity.cu).  q is [n][21] of float or double on the device.
cudaError_t launch_validity(const ValidityTables& tab, const void* q, int dtype, int64_t n, uint8_t* valid,
                            uint8_t* reason, float* margins, bool collision_only, bool static_pairs, cudaStream_t stream,
                            const GatherPeers* gather = nullptr);
cudaError_t launch_fk(const void* q, int dtype, int64_t n, double* poses, double* com, cudaStream_t stream);
cudaError_t launch_ffma_peak(int sm_count, int iters, float* sink, cudaStream_t stream, double* flops_out);

// Pooled forest of the RRT driver: 2 trees per query, fixed capacity each (device pointers).
//   soa float [T][21][cap] | aos double [T][cap][21] | pred int32 [T][cap] | size int32 [T]
//   hand int32 [T][cap]: Node::cart_hand_pose_index (RPh:35), only with an articulation constraint
struct ForestView {
  float* soa;
  double* aos;
  int32_t* pred;
  int32_t* size;
  int64_t cap;
  int32_t* hand;
};
// Per-query control state of the lock-step RRT-Connect driver (device arrays of length nq).
struct RrtState {
  uint8_t* active;      // query still searching
  uint8_t* swap;        // 0: extend the start tree this iteration, 1: the goal tree (RP:1073,1136)
  uint8_t* walking;     // the extend step succeeded: the other tree runs connectTree this iteration
  int32_t* iterations;
  int32_t* success;
  int32_t* connector;   // 1: goal tree connected to start tree, 2: the reverse (RP:1101,1161)
  int32_t* bridge;      // q_near.index of the connecting walk (RP:1000-1001)
  int32_t* tree_a;      // tree extended this iteration
  int32_t* tree_b;      // tree that tries to connect
  int32_t* idx_b;       // nearest neighbour of q_connect in tree_b
  double* q_connect;    // node just added to tree_a
  int32_t* connect_hand;// its cart_hand_pose_index (articulation constraint only)
  double* cstart;       // configuration of that nearest neighbour
  int32_t* n_active;    // [1]
  int32_t* overflow;    // [1] a tree outgrew its capacity
};

// planner kernels (planner_kernels.cu); all pointers are device pointers
struct PlanConst;
// ma_pts: [nq][n_pts][6] hand trajectory per query (device) or nullptr = no articulation constraint
cudaError_t launch_connect_forest(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, ForestView F, RrtState S,
                                  int nq, const double* ma_pts, int n_pts, cudaStream_t s);
// the whole solveQuery loop of nq queries in one launch (one warp per query, queries finish independently)
cudaError_t launch_rrt_persistent(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, ForestView F, RrtState S,
                                  const double* db, int n_db, const uint64_t* seeds, int max_iter, int nq, const double* ma_pts,
                                  int n_pts, cudaStream_t s);
// roots validated + inserted, loop, results and raw paths in one launch (planner_kernels.cu: rrt_fused_kernel)
cudaError_t launch_rrt_fused(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, ForestView F, const double* starts,
                             const double* goals, const uint64_t* seeds, int32_t* res, double* path_rows,
                             unsigned long long* path_used, long long path_cap, const double* db, int n_db, int max_iter, int nq,
                             const double* ma_pts, int n_pts, int32_t* overflow_flag, cudaStream_t s);
// scratch: 256 + 4 n bytes of device memory selects the two-launch form for n >= 4096 (nullptr: one launch)
size_t steer_scratch_bytes(int64_t n);     // scratch of the phased form of launch_steer_project (n >= 4096)
cudaError_t launch_steer_project(const PlanConst& pc, const double* q_near, const double* q_target, int64_t n, double* q_out,
                                 int32_t* status, int32_t* ds_status, int32_t* iters, void* scratch, cudaStream_t s);
// rows [e][0 .. n_added[e]) of nodes [n_edges][max_steps][21] -> packed[offset[e] ..] (dense, edge order)
cudaError_t launch_pack_rows(const double* nodes, const int32_t* n_added, const int64_t* offset, int64_t n_edges, int max_steps,
                             double* packed, cudaStream_t s);
cudaError_t launch_connect(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, const double* q_start,
                           const double* q_connect, int64_t n_edges, int max_steps, double* nodes_out, int32_t* n_added,
                           int32_t* status, cudaStream_t s);
cudaError_t launch_edge_check(const ValidityTables& tab, bool static_pairs, const double* qa, const double* qb, int64_t n_edges,
                              int k, uint8_t* ok, int32_t* first_bad, cudaStream_t s);
// scratch: sampler_scratch_bytes(count) bytes of device memory (work lists of the phased sampler, planner_kernels.cu)
size_t sampler_scratch_bytes(int64_t count);
cudaError_t launch_sample(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, uint64_t seed, uint64_t first,
                          int64_t count, int max_ik_iter, double* rows, uint64_t* row_index, int64_t cap,
                          unsigned long long* n_out, uint8_t* stage_out, void* scratch, cudaStream_t s);


// right-arm kernels (arm_kernels.cu)
cudaError_t launch_arm_ik(const double* targets, int64_t n, double* sols, int32_t* counts, cudaStream_t s);
cudaError_t launch_goal_arm_ik(const PlanConst& pc, const double* pose6, uint64_t seed, uint64_t first_index, const double* q,
                               int64_t n, uint8_t* ok, double* arm, double* erg, cudaStream_t s);
cudaError_t launch_enforce_ma(const PlanConst& pc, const double* pts_dev, int n_pts, int start_tree, const int32_t* near_hand,
                              const double* q_near, const double* q_ds, int64_t n, double* q_out, int32_t* status, cudaStream_t s);
// nq requests (poses_dev [nq][6], seeds_dev [nq]) x count sample indices; compacted outputs are per request:
// rows [nq][cap][21], row_index / row_erg [nq][cap], n_out [nq]; the per-sample outputs [nq * count] are optional.
// With row_req the requests share one pool instead: rows [cap][21], row_index / row_erg / row_req [cap], n_out [1].
cudaError_t launch_goal_sample(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, const double* poses_dev,
                               const uint64_t* seeds_dev, int nq, uint64_t first, int64_t count, int max_ik_iter, double* rows,
                               uint64_t* row_index, double* row_erg, int64_t cap, unsigned long long* n_out, uint8_t* stage_out,
                               double* erg_all, double* rows_all, int32_t* row_req, cudaStream_t s);
// the same sampler without per-sample outputs, as three launches over densely packed work lists (arm_kernels.cu);
// scratch: goal_sample_phased_scratch_bytes(nq * count) bytes of device memory
size_t goal_sample_phased_scratch_bytes(int64_t items);
cudaError_t launch_goal_sample_phased(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, const double* poses_dev,
                                      const uint64_t* seeds_dev, int nq, uint64_t first, int64_t count, int max_ik_iter, double* rows,
                                      uint64_t* row_index, double* row_erg, int64_t cap, unsigned long long* n_out, int32_t* row_req,
                                      void* scratch, cudaStream_t s);
cudaError_t launch_ma_project(const PlanConst& pc, const double* ma_pts, int n_pts, ForestView F, RrtState S, const int32_t* idx_a,
                              const double* q_near, double* q_new, int32_t* status, int nq, cudaStream_t s);

}  // namespace nwb

struct nwb_ctx : nwb::Ctx {};

// ---- helpers shared by the extern "C" translation units ----------------------------------------------------------------
// CUDA call -> NWB_E_CUDA with the failing expression and the runtime's message in nwb_last_error()
#define NWB_TRY_CUDA(ctx, call)                                                \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) {                                                   \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);         \
      return NWB_E_CUDA;                                                       \
    }                                                                          \
  } while (0)

namespace nwb {
// bump allocator over the context's planner scratch buffer (grow-only), 256-byte aligned sections
struct DeviceArena {
  nwb_ctx* ctx;
  size_t off = 0;
  explicit DeviceArena(nwb_ctx* c) : ctx(c) {}
  size_t add(size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  }
  int commit() {
    if (off <= ctx->d_plan_bytes) return NWB_OK;
    if (ctx->d_plan) cudaFree(ctx->d_plan);
    ctx->d_plan = nullptr;
    ctx->d_plan_bytes = 0;
    size_t cap = off > ((size_t)1 << 20) ? off : ((size_t)1 << 20);
    cudaError_t e = cudaMalloc(&ctx->d_plan, cap);
    if (e != cudaSuccess) {
      ctx->err = std::string("cudaMalloc(scratch): ") + cudaGetErrorString(e);
      return NWB_E_CUDA;
    }
    ctx->d_plan_bytes = cap;
    return NWB_OK;
  }
  template <class T>
  T* at(size_t o) const { return reinterpret_cast<T*>(static_cast<char*>(ctx->d_plan) + o); }
};
}  // namespace nwb
