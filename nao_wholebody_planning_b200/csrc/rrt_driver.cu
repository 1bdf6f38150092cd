// rrt_driver.cu - host side of the RRT-Connect driver (device loop: rrt_persistent_kernel in planner_kernels.cu).
//
// Replaces RRT_Planner::setStartGoalConfigs / solveQuery / extendTree / connectTree /
// addConfigtoTree / writePath (rrt_connect_planner/src/rrt_planner.cpp:227-302, 1039-1212, 787-884,
// 890-1034, 749-783, 1303-1392) and path_shortcutter / interpolateShortPathWaypoints (RP:1830-1994)
// for nq INDEPENDENT planning queries advanced in lock step: iteration i of every still-active
// query is one pass of
//     random DB row -> nearest neighbour -> steer -> double-support projection -> validity -> append
//     -> nearest neighbour in the other tree -> connect walk -> append -> swap
// where each arrow is one kernel over all queries.  The control state (which queries are active,
// which tree extends, success / bridge nodes) lives on the device too, so an iteration is seven
// kernel launches with NO host synchronisation; the host only reads the active-query count every
// few iterations.  A single query is the nq = 1 case (latency bound: RRT-Connect is serial by construction,
// SURVEY.md section 8e "a single RRT query: replicas only").
//
// Trees live in one pooled forest (2 trees per query, fixed capacity each):
//   soa float [T][21][C] | aos double [T][C][21] | pred int32 [T][C] | size int32 [T]
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "arm_core.cuh"
#include "nwb_internal.h"
#include "planner_core.cuh"

namespace nwb {
void make_plan_const(const nwb_ctx* ctx, const double* weights, double step, PlanConst& pc);

__device__ __forceinline__ bool lex_less2(double da, int ia, double db, int ib) {
  return (da < db) || (da == db && (unsigned)ia < (unsigned)ib);
}

// start of iteration `iteration` for every active query: which tree extends (RP:1073,1136), and
// q_rand[k] = DB row floor(U * (n-1)) (quirk C3) or floor(U * n) with U = Philox(seed_k; iteration)
// (getRandomStableConfig, RP:316-332, with the injected stream)
__global__ void rrt_begin_kernel(const double* __restrict__ db, int n_db, const uint64_t* __restrict__ seeds, int iteration,
                                 unsigned quirks, int nq, RrtState S, double* __restrict__ q_rand) {
  int k = blockIdx.x;
  if (k >= nq || !S.active[k]) return;
  __shared__ int row;
  if (threadIdx.x == 0) {
    uint32_t r[4];
    uint64_t s = seeds[k];
    philox4x32_10((uint32_t)iteration, 0u, 0u, kStreamPlanner, (uint32_t)s, (uint32_t)(s >> 32), r);
    double u = u01(r[0]);
    row = (quirks & NWB_QUIRK_DB_INDEX_FLOOR) ? (int)floor(0.0 + (double)(n_db - 1 - 0) * u) : (int)floor((double)n_db * u);
    const int sw = S.swap[k];
    S.tree_a[k] = 2 * k + sw;
    S.tree_b[k] = 2 * k + 1 - sw;
    S.iterations[k] = iteration + 1;
    S.walking[k] = 0;
  }
  __syncthreads();
  if (threadIdx.x < NWB_NJ) q_rand[k * NWB_NJ + threadIdx.x] = db[(size_t)row * NWB_NJ + threadIdx.x];
}

// one block per query: exact nearest neighbour of q[k] in tree tree_of[k]; also gathers that node
__global__ void __launch_bounds__(128)
forest_nn_kernel(ForestView F, const int32_t* __restrict__ tree_of, const uint8_t* __restrict__ active,
                 const double* __restrict__ q, int nq, int32_t* __restrict__ idx_out, double* __restrict__ near_out) {
  const int k = blockIdx.x;
  if (k >= nq || (active && !active[k])) return;
  const int t = tree_of[k];
  const int N = F.size[t];
  const float* soa = F.soa + (size_t)t * NWB_NJ * F.cap;
  const double* aos = F.aos + (size_t)t * F.cap * NWB_NJ;
  __shared__ double qd[NWB_NJ];
  __shared__ float qf[NWB_NJ];
  __shared__ double red_d[4];
  __shared__ int red_i[4];
  if (threadIdx.x < NWB_NJ) {
    double v = q[(size_t)k * NWB_NJ + threadIdx.x];
    qd[threadIdx.x] = v;
    qf[threadIdx.x] = (float)v;
  }
  __syncthreads();
  double best = 10000.0;
  int besti = -1;
  float thr = 3.0e38f;
  for (int node = threadIdx.x; node < N; node += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) {
      float d = qf[j] - soa[(size_t)j * F.cap + node];
      s = fmaf(d, d, s);
    }
    if (s <= thr) {
      double sum = 0.0;
      const double* row = aos + (size_t)node * NWB_NJ;
      for (int j = 0; j < NWB_NJ; ++j) {
        double d = __dsub_rn(qd[j], row[j]);
        sum = __dadd_rn(sum, __dmul_rn(d, d));
      }
      double dn = sqrt(sum);
      if (dn < best) { best = dn; besti = node; }
      thr = __double2float_ru((best * best + 1.2e-5 * best + 1e-10) * (1.0 + 1e-6));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double d2 = __shfl_xor_sync(0xffffffffu, best, o);
    int i2 = __shfl_xor_sync(0xffffffffu, besti, o);
    if (lex_less2(d2, i2, best, besti)) { best = d2; besti = i2; }
  }
  if ((threadIdx.x & 31) == 0) { red_d[threadIdx.x >> 5] = best; red_i[threadIdx.x >> 5] = besti; }
  __syncthreads();
  __shared__ int winner;
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (lex_less2(red_d[w], red_i[w], best, besti)) { best = red_d[w]; besti = red_i[w]; }
    idx_out[k] = besti;
    winner = besti;
  }
  __syncthreads();
  if (threadIdx.x < NWB_NJ && winner >= 0) near_out[(size_t)k * NWB_NJ + threadIdx.x] = aos[(size_t)winner * NWB_NJ + threadIdx.x];
}

// append count[k] nodes (src[k][stride][21]) to tree tree_of[k]; first predecessor = pred_first[k],
// later ones chain to the node appended just before.  first_index[k] = index of the first appended node.
__global__ void forest_append_kernel(ForestView F, const int32_t* __restrict__ tree_of, const double* __restrict__ src, int stride,
                                     const int32_t* __restrict__ count, const int32_t* __restrict__ pred_first, int nq,
                                     int32_t* __restrict__ first_index, int32_t* __restrict__ overflow) {
  const int k = blockIdx.x;
  if (k >= nq) return;
  const int n = count[k];
  if (n <= 0) return;
  const int t = tree_of[k];
  const int base = F.size[t];
  if (base + n > F.cap) {
    if (threadIdx.x == 0) atomicExch(overflow, 1);
    return;
  }
  float* soa = F.soa + (size_t)t * NWB_NJ * F.cap;
  double* aos = F.aos + (size_t)t * F.cap * NWB_NJ;
  for (int e = threadIdx.x; e < n * NWB_NJ; e += blockDim.x) {
    int i = e / NWB_NJ, j = e % NWB_NJ;
    double v = src[((size_t)k * stride + i) * NWB_NJ + j];
    aos[(size_t)(base + i) * NWB_NJ + j] = v;
    soa[(size_t)j * F.cap + base + i] = (float)v;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) F.pred[(size_t)t * F.cap + base + i] = (i == 0) ? pred_first[k] : base + i - 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    F.size[t] = base + n;
    if (first_index) first_index[k] = base;
  }
}

// after steer/project + validity: extendTree's decision (RP:846-870).  A surviving q_new is appended to
// tree_a (predecessor = its nearest neighbour) and becomes q_connect of this iteration's connectTree.
__global__ void rrt_extend_commit_kernel(ForestView F, RrtState S, const int32_t* __restrict__ status, const uint8_t* __restrict__ valid,
                                         const int32_t* __restrict__ idx_a, const double* __restrict__ q_new, int nq, int n_pts) {
  const int k = blockIdx.x;
  if (k >= nq || !S.active[k]) return;
  if (status[k] == NWB_TRAPPED || !valid[k]) return;        // walking stays 0: only the swap happens this iteration
  const int t = S.tree_a[k];
  const int base = F.size[t];
  if (base >= F.cap) {
    if (threadIdx.x == 0) {
      atomicExch(S.overflow, 1);
      S.active[k] = 0;
      atomicSub(S.n_active, 1);
    }
    return;
  }
  if (threadIdx.x < NWB_NJ) {
    const double v = q_new[(size_t)k * NWB_NJ + threadIdx.x];
    F.aos[((size_t)t * F.cap + base) * NWB_NJ + threadIdx.x] = v;
    F.soa[(size_t)t * NWB_NJ * F.cap + (size_t)threadIdx.x * F.cap + base] = (float)v;
    S.q_connect[(size_t)k * NWB_NJ + threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    F.pred[(size_t)t * F.cap + base] = idx_a[k];
    F.size[t] = base + 1;
    S.walking[k] = 1;
    if (n_pts > 0) {                                        // addConfigtoTree's hand index (RP:757-767)
      const int h = next_hand_index(MaView{nullptr, n_pts}, (t & 1) == 0, F.hand[(size_t)t * F.cap + idx_a[k]]);
      F.hand[(size_t)t * F.cap + base] = h;
      S.connect_hand[k] = h;
    }
  }
}

// setStartGoalConfigs with an active articulation constraint (RP:274-282): the start root sits on the
// first trajectory point, the goal root on the last.
__global__ void forest_root_hand_kernel(ForestView F, int n_trees, int n_pts) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_trees) F.hand[(size_t)t * F.cap] = (t & 1) ? n_pts - 1 : 0;
}

// writePath, path extraction part (RP:1303-1392), on the device.  Pass 1 (rows == nullptr): length of the start-
// and goal-tree branches of every solved query.  Pass 2: the way-points, start root ... bridge | bridge ... goal
// root, packed at offset[k].  One warp per query; the predecessor walk is serial, the row copies use the lanes.
__global__ void __launch_bounds__(32)
rrt_path_kernel(ForestView F, const int32_t* __restrict__ success, const int32_t* __restrict__ connector,
                const int32_t* __restrict__ bridge, int nq, int32_t* __restrict__ lens /*[nq][2]*/,
                const int64_t* __restrict__ offset, double* __restrict__ rows) {
  const int k = blockIdx.x, lane = threadIdx.x;
  if (k >= nq || !success[k]) return;
  const int ts = 2 * k, tg = 2 * k + 1;
  int idx_s, idx_g;
  if (connector[k] == 1) { idx_s = F.size[ts] - 1; idx_g = bridge[k]; }       // RP:1317-1327
  else { idx_s = bridge[k]; idx_g = F.size[tg] - 1; }
  const int32_t* pred_s = F.pred + (size_t)ts * F.cap;
  const int32_t* pred_g = F.pred + (size_t)tg * F.cap;
  if (!rows) {
    if (lane == 0) {
      int ls = 1, lg = 1;
      for (int i = idx_s; i > 0; i = pred_s[i]) ++ls;
      for (int i = idx_g; i > 0; i = pred_g[i]) ++lg;
      lens[2 * k] = ls;
      lens[2 * k + 1] = lg;
    }
    return;
  }
  const int ls = lens[2 * k], lg = lens[2 * k + 1];
  double* out = rows + offset[k] * NWB_NJ;
  int i = idx_s;
  for (int pos = ls - 1; pos >= 0; --pos) {                                    // start branch, reversed into place
    if (lane < NWB_NJ) out[(size_t)pos * NWB_NJ + lane] = F.aos[((size_t)ts * F.cap + i) * NWB_NJ + lane];
    i = pred_s[i];
  }
  i = idx_g;
  for (int pos = 0; pos < lg; ++pos) {                                         // goal branch as walked
    if (lane < NWB_NJ) out[(size_t)(ls + pos) * NWB_NJ + lane] = F.aos[((size_t)tg * F.cap + i) * NWB_NJ + lane];
    i = pred_g[i];
  }
}

}  // namespace nwb

using namespace nwb;

namespace {
#define R_CUDA(ctx, call) NWB_TRY_CUDA(ctx, call)

}  // namespace

// One upload, one launch, one download: the default form of solveQuery for nq queries (rrt_fused_kernel).
static int solve_fused(nwb_ctx* ctx, const double* starts, const double* goals, int64_t nq, const double* db, int32_t n_db,
                       const double* weights, const uint64_t* seeds, int32_t max_iter, double step, int32_t tree_capacity,
                       const double* hand_traj, int P, nwb_plan_result* results, double* raw_paths, int32_t raw_path_capacity) {
  cudaStream_t s = ctx->stream;
  const int Q = (int)nq, T = 2 * Q;
  const int64_t C = tree_capacity > 0 ? tree_capacity : 4096;
  const bool sp = ctx->params.srdf_trim != 0 && !ctx->force_table_pairs;
  PlanConst pc;
  make_plan_const(ctx, weights, step, pc);
  auto grow = [&](void** buf, size_t* have, size_t need) -> cudaError_t {
    if (*have >= need) return cudaSuccess;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc(buf, need);
    if (e == cudaSuccess) *have = need;
    return e;
  };
  const size_t soa_b = sizeof(float) * NWB_NJ * C * T, aos_b = sizeof(double) * NWB_NJ * C * T, pred_b = sizeof(int32_t) * C * T;
  const size_t hand_b = P ? pred_b : 0;
  R_CUDA(ctx, grow(&ctx->d_forest, &ctx->d_forest_bytes, soa_b + aos_b + pred_b + hand_b + 256));
  char* FB = static_cast<char*>(ctx->d_forest);

  // the stable-configuration database of the services stays on the device between calls
  const double* d_db = nullptr;
  bool db_inline = true;
  if (db == ctx->ds_db.data() && (size_t)n_db * NWB_NJ == ctx->ds_db.size()) {
    if (ctx->d_dbcache_version != ctx->ds_db_version || ctx->d_dbcache_rows != (size_t)n_db) {
      if (ctx->d_dbcache) cudaFree(ctx->d_dbcache);
      ctx->d_dbcache = nullptr;
      R_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_dbcache), sizeof(double) * NWB_NJ * (size_t)n_db));
      R_CUDA(ctx, cudaMemcpyAsync(ctx->d_dbcache, db, sizeof(double) * NWB_NJ * (size_t)n_db, cudaMemcpyHostToDevice, s));
      ctx->d_dbcache_rows = (size_t)n_db;
      ctx->d_dbcache_version = ctx->ds_db_version;
    }
    d_db = ctx->d_dbcache;
    db_inline = false;
  }

  // device arena: [inputs: starts | goals | seeds | hand trajectories | database] [outputs: res | used | overflow | paths]
  const size_t qbytes = sizeof(double) * NWB_NJ * Q;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_size = carve(sizeof(int32_t) * T);
  const size_t o_in = off;
  const size_t o_starts = carve(qbytes), o_goals = carve(qbytes), o_seeds = carve(sizeof(uint64_t) * Q),
               o_ma = carve(sizeof(double) * 6 * (size_t)P * Q), o_db = carve(db_inline ? sizeof(double) * NWB_NJ * (size_t)n_db : 0);
  const size_t in_bytes = off - o_in;
  const size_t o_out = off;
  const size_t o_res = carve(sizeof(int32_t) * 12 * Q), o_used = carve(16), o_ovf = carve(16);
  const long long path_cap = std::max<long long>((long long)Q * 512, 2 * C);   // one query's path always fits: <= 2 C rows
  const size_t o_paths = carve(sizeof(double) * NWB_NJ * (size_t)path_cap);
  const size_t head_bytes = o_paths - o_out;                             // res + used + overflow
  R_CUDA(ctx, grow(&ctx->d_rrt, &ctx->d_rrt_bytes, off));
  char* M = static_cast<char*>(ctx->d_rrt);
  ForestView F{reinterpret_cast<float*>(FB), reinterpret_cast<double*>(FB + ((soa_b + 255) & ~(size_t)255)),
               reinterpret_cast<int32_t*>(FB + ((soa_b + 255) & ~(size_t)255) + aos_b), reinterpret_cast<int32_t*>(M + o_size), C,
               P ? reinterpret_cast<int32_t*>(FB + ((soa_b + 255) & ~(size_t)255) + aos_b + pred_b) : nullptr};

  // pinned staging: inputs in the arena's layout, so one copy moves them all; results come back the same way
  const int64_t spec_rows = Q <= 4 ? std::min<long long>(path_cap, 768) : 0;   // small calls: fetch the path rows speculatively
  const size_t back_bytes = head_bytes + sizeof(double) * NWB_NJ * (size_t)spec_rows;
  const size_t pin_need = std::max(in_bytes, back_bytes) + 256;
  if (ctx->h_pin_bytes < pin_need) {
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    ctx->h_pin = nullptr;
    ctx->h_pin_bytes = 0;
    R_CUDA(ctx, cudaMallocHost(&ctx->h_pin, std::max(pin_need, (size_t)1 << 20)));
    ctx->h_pin_bytes = std::max(pin_need, (size_t)1 << 20);
  }
  char* H = static_cast<char*>(ctx->h_pin);
  std::memcpy(H + (o_starts - o_in), starts, qbytes);
  std::memcpy(H + (o_goals - o_in), goals, qbytes);
  std::memcpy(H + (o_seeds - o_in), seeds, sizeof(uint64_t) * Q);
  if (P) std::memcpy(H + (o_ma - o_in), hand_traj, sizeof(double) * 6 * (size_t)P * Q);
  if (db_inline) std::memcpy(H + (o_db - o_in), db, sizeof(double) * NWB_NJ * (size_t)n_db);
  R_CUDA(ctx, cudaMemcpyAsync(M + o_in, H, in_bytes, cudaMemcpyHostToDevice, s));
  R_CUDA(ctx, cudaMemsetAsync(M + o_used, 0, o_paths - o_used, s));       // row counter + overflow flag
  if (db_inline) d_db = reinterpret_cast<const double*>(M + o_db);
  R_CUDA(ctx, launch_rrt_fused(pc, ctx->tab_group, sp, F, reinterpret_cast<const double*>(M + o_starts),
                               reinterpret_cast<const double*>(M + o_goals), reinterpret_cast<const uint64_t*>(M + o_seeds),
                               reinterpret_cast<int32_t*>(M + o_res), reinterpret_cast<double*>(M + o_paths),
                               reinterpret_cast<unsigned long long*>(M + o_used), path_cap, d_db, n_db, max_iter, Q,
                               P ? reinterpret_cast<const double*>(M + o_ma) : nullptr, P, reinterpret_cast<int32_t*>(M + o_ovf), s));
  R_CUDA(ctx, cudaMemcpyAsync(H, M + o_out, back_bytes, cudaMemcpyDeviceToHost, s));
  R_CUDA(ctx, cudaStreamSynchronize(s));
  int64_t launches = 1;

  const int32_t* res = reinterpret_cast<const int32_t*>(H + (o_res - o_out));
  const unsigned long long used = *reinterpret_cast<const unsigned long long*>(H + (o_used - o_out));
  const int32_t overflow = *reinterpret_cast<const int32_t*>(H + (o_ovf - o_out));
  // A query whose tree outgrew the capacity, or whose raw path did not fit the shared row pool, is NOT an error of the
  // call: the reference's trees are unbounded std::vectors (RP:1067-1198).  Such queries are solved again on their own
  // with doubled capacity (same seed, hence the same search) while every other result of the batch stands.
  std::vector<int> redo;
  for (int k = 0; k < Q; ++k) {
    const int32_t* r = res + (size_t)k * 12;
    nwb_plan_result& o = results[k];
    std::memset(&o, 0, sizeof o);
    o.start_valid = r[0];
    o.goal_valid = r[1];
    o.success = r[2];
    o.iterations = r[3];
    o.connector = r[4];
    o.bridge_other = r[5];
    o.nodes_start = r[6];
    o.nodes_goal = r[7];
    o.num_nodes_generated = o.success ? o.nodes_start + o.nodes_goal : 0;   // RP:1045: stays 0 unless a path is found
    o.raw_path_len = o.success ? r[8] : 0;
    if (r[10] || (o.success && r[9] < 0)) redo.push_back(k);
  }
  (void)overflow;
  if (raw_paths && used > 0) {
    const double* rows = reinterpret_cast<const double*>(H + head_bytes);
    std::vector<double> big;
    if ((int64_t)used > spec_rows) {                                       // fetch exactly what was written
      big.resize((size_t)used * NWB_NJ);
      R_CUDA(ctx, cudaMemcpyAsync(big.data(), M + o_paths, sizeof(double) * NWB_NJ * (size_t)used, cudaMemcpyDeviceToHost, s));
      R_CUDA(ctx, cudaStreamSynchronize(s));
      rows = big.data();
    }
    for (int k = 0; k < Q; ++k) {
      const int32_t* r = res + (size_t)k * 12;
      if (!r[2] || r[9] < 0 || r[10]) continue;
      const int n = std::min<int>(r[8], raw_path_capacity);                // the caller sees the full length in raw_path_len
      std::memcpy(raw_paths + (size_t)k * raw_path_capacity * NWB_NJ, rows + (size_t)r[9] * NWB_NJ, sizeof(double) * NWB_NJ * n);
    }
  }
  ctx->last_launches = launches;
  if (!redo.empty()) {
    constexpr int64_t kMaxTreeCapacity = (int64_t)1 << 21;                  // 2 M nodes per tree: 1.1 GB per query
    if (2 * C > kMaxTreeCapacity) {
      for (int k : redo) {
        results[k].success = 0;
        results[k].raw_path_len = 0;
        results[k].num_nodes_generated = 0;
      }
      ctx->err = "nwb_solve_query_batch: a tree outgrew the largest supported capacity (2^21 nodes); that query is reported as failed";
      return NWB_E_NOMEM;
    }
    const int R = (int)redo.size();
    std::vector<double> st((size_t)R * NWB_NJ), go((size_t)R * NWB_NJ), ht(P ? (size_t)R * 6 * P : 0);
    std::vector<uint64_t> sd(R);
    std::vector<nwb_plan_result> rr(R);
    std::vector<double> rp(raw_paths ? (size_t)R * raw_path_capacity * NWB_NJ : 0);
    for (int i = 0; i < R; ++i) {
      const int k = redo[i];
      std::memcpy(&st[(size_t)i * NWB_NJ], starts + (size_t)k * NWB_NJ, sizeof(double) * NWB_NJ);
      std::memcpy(&go[(size_t)i * NWB_NJ], goals + (size_t)k * NWB_NJ, sizeof(double) * NWB_NJ);
      if (P) std::memcpy(&ht[(size_t)i * 6 * P], hand_traj + (size_t)k * 6 * P, sizeof(double) * 6 * P);
      sd[i] = seeds[k];
    }
    const int64_t launches_so_far = ctx->last_launches;
    int rc = NWB_OK;
    for (int i = 0; i < R && rc == NWB_OK; ++i)                            // one at a time: the memory goes into ONE big forest
      rc = solve_fused(ctx, &st[(size_t)i * NWB_NJ], &go[(size_t)i * NWB_NJ], 1, db, n_db, weights, &sd[i], max_iter, step,
                       (int32_t)(2 * C), P ? &ht[(size_t)i * 6 * P] : nullptr, P, &rr[i],
                       raw_paths ? &rp[(size_t)i * raw_path_capacity * NWB_NJ] : nullptr, raw_path_capacity);
    for (int i = 0; i < R; ++i) {
      const int k = redo[i];
      results[k] = rr[i];
      if (raw_paths && rr[i].success)
        std::memcpy(raw_paths + (size_t)k * raw_path_capacity * NWB_NJ, &rp[(size_t)i * raw_path_capacity * NWB_NJ],
                    sizeof(double) * NWB_NJ * std::min<int>(rr[i].raw_path_len, raw_path_capacity));
    }
    ctx->last_launches += launches_so_far;
    return rc;
  }
  return NWB_OK;
}

static int solve_impl(nwb_ctx* ctx, const double* starts, const double* goals, int64_t nq, const double* db, int32_t n_db,
                      const double* weights, const uint64_t* seeds, int32_t max_iter, double step, int32_t tree_capacity,
                      const double* hand_traj, int32_t n_pts, nwb_plan_result* results, double* raw_paths,
                      int32_t raw_path_capacity) {
  if (!ctx) return NWB_E_INVALID;
  if (nq < 0 || (nq > 0 && (!starts || !goals || !db || !seeds || !results)) || n_db < 2 || max_iter < 0 || !(step > 0) ||
      (raw_paths && raw_path_capacity <= 0) || (hand_traj && n_pts < 2)) {
    ctx->err = "nwb_solve_query_batch: bad argument";
    return NWB_E_INVALID;
  }
  const int P = hand_traj ? n_pts : 0;
  if (nq == 0) return NWB_OK;
  R_CUDA(ctx, cudaSetDevice(ctx->device));
  // default: the fused form; NWB_RRT_STAGED=1 (separate root validation / path extraction around the one-launch loop)
  // and NWB_RRT_LOCKSTEP=1 (7-8 kernels per iteration) keep the earlier drivers for comparison
  static const bool staged = getenv("NWB_RRT_STAGED") != nullptr || getenv("NWB_RRT_LOCKSTEP") != nullptr;
  if (!staged)
    return solve_fused(ctx, starts, goals, nq, db, n_db, weights, seeds, max_iter, step, tree_capacity, hand_traj, P, results,
                       raw_paths, raw_path_capacity);
  cudaStream_t s = ctx->stream;
  const int Q = (int)nq, T = 2 * Q;
  const int64_t C = tree_capacity > 0 ? tree_capacity : 4096;
  const bool sp = ctx->params.srdf_trim != 0 && !ctx->force_table_pairs;
  PlanConst pc;
  make_plan_const(ctx, weights, step, pc);

  // one device allocation for the forest, one for everything else
  const size_t qbytes = sizeof(double) * NWB_NJ * Q;
  auto grow = [&](void** buf, size_t* have, size_t need) -> cudaError_t {
    if (*have >= need) return cudaSuccess;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc(buf, need);
    if (e == cudaSuccess) *have = need;
    return e;
  };
  const size_t soa_b = sizeof(float) * NWB_NJ * C * T, aos_b = sizeof(double) * NWB_NJ * C * T, pred_b = sizeof(int32_t) * C * T;
  const size_t hand_b = P ? pred_b : 0;
  R_CUDA(ctx, grow(&ctx->d_forest, &ctx->d_forest_bytes, soa_b + aos_b + pred_b + hand_b + 256));
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_size = carve(sizeof(int32_t) * T), o_db = carve(sizeof(double) * NWB_NJ * n_db), o_seeds = carve(sizeof(uint64_t) * Q),
               o_sg = carve(2 * qbytes), o_qrand = carve(qbytes), o_near = carve(qbytes), o_qnew = carve(qbytes),
               o_qconn = carve(qbytes), o_cstart = carve(qbytes), o_status = carve(sizeof(int32_t) * Q), o_valid = carve(2 * (size_t)Q),
               o_active = carve(Q), o_swap = carve(Q), o_walk = carve(Q), o_iter = carve(sizeof(int32_t) * Q),
               o_succ = carve(sizeof(int32_t) * Q), o_conn = carve(sizeof(int32_t) * Q), o_bridge = carve(sizeof(int32_t) * Q),
               o_ta = carve(sizeof(int32_t) * Q), o_tb = carve(sizeof(int32_t) * Q), o_ia = carve(sizeof(int32_t) * Q),
               o_ib = carve(sizeof(int32_t) * Q), o_cnt = carve(sizeof(int32_t) * T), o_roots = carve(sizeof(int32_t) * T),
               o_zero = carve(sizeof(int32_t) * T), o_nact = carve(16), o_ovf = carve(16), o_chand = carve(sizeof(int32_t) * Q),
               o_ma = carve(sizeof(double) * 6 * (size_t)P * Q);
  R_CUDA(ctx, grow(&ctx->d_rrt, &ctx->d_rrt_bytes, off));
  R_CUDA(ctx, cudaMemsetAsync(ctx->d_rrt, 0, off, s));           // inactive slots still flow through the kernels: keep them finite
  char* M = static_cast<char*>(ctx->d_rrt);
  char* FB = static_cast<char*>(ctx->d_forest);
  ForestView F{reinterpret_cast<float*>(FB), reinterpret_cast<double*>(FB + ((soa_b + 255) & ~(size_t)255)),
               reinterpret_cast<int32_t*>(FB + ((soa_b + 255) & ~(size_t)255) + aos_b), reinterpret_cast<int32_t*>(M + o_size), C,
               P ? reinterpret_cast<int32_t*>(FB + ((soa_b + 255) & ~(size_t)255) + aos_b + pred_b) : nullptr};
  RrtState S;
  S.active = reinterpret_cast<uint8_t*>(M + o_active);
  S.swap = reinterpret_cast<uint8_t*>(M + o_swap);
  S.walking = reinterpret_cast<uint8_t*>(M + o_walk);
  S.iterations = reinterpret_cast<int32_t*>(M + o_iter);
  S.success = reinterpret_cast<int32_t*>(M + o_succ);
  S.connector = reinterpret_cast<int32_t*>(M + o_conn);
  S.bridge = reinterpret_cast<int32_t*>(M + o_bridge);
  S.tree_a = reinterpret_cast<int32_t*>(M + o_ta);
  S.tree_b = reinterpret_cast<int32_t*>(M + o_tb);
  S.idx_b = reinterpret_cast<int32_t*>(M + o_ib);
  S.q_connect = reinterpret_cast<double*>(M + o_qconn);
  S.cstart = reinterpret_cast<double*>(M + o_cstart);
  S.n_active = reinterpret_cast<int32_t*>(M + o_nact);
  S.overflow = reinterpret_cast<int32_t*>(M + o_ovf);
  S.connect_hand = reinterpret_cast<int32_t*>(M + o_chand);
  double* d_ma = P ? reinterpret_cast<double*>(M + o_ma) : nullptr;
  double* d_db = reinterpret_cast<double*>(M + o_db);
  uint64_t* d_seeds = reinterpret_cast<uint64_t*>(M + o_seeds);
  double* d_sg = reinterpret_cast<double*>(M + o_sg);
  double* d_qrand = reinterpret_cast<double*>(M + o_qrand);
  double* d_near = reinterpret_cast<double*>(M + o_near);
  double* d_qnew = reinterpret_cast<double*>(M + o_qnew);
  int32_t* d_status = reinterpret_cast<int32_t*>(M + o_status);
  uint8_t* d_valid = reinterpret_cast<uint8_t*>(M + o_valid);
  int32_t* d_idx_a = reinterpret_cast<int32_t*>(M + o_ia);

  R_CUDA(ctx, cudaMemcpyAsync(d_db, db, sizeof(double) * NWB_NJ * n_db, cudaMemcpyHostToDevice, s));
  R_CUDA(ctx, cudaMemcpyAsync(d_seeds, seeds, sizeof(uint64_t) * Q, cudaMemcpyHostToDevice, s));
  if (P) R_CUDA(ctx, cudaMemcpyAsync(d_ma, hand_traj, sizeof(double) * 6 * (size_t)P * Q, cudaMemcpyHostToDevice, s));
  R_CUDA(ctx, cudaMemcpyAsync(d_sg, starts, qbytes, cudaMemcpyHostToDevice, s));
  R_CUDA(ctx, cudaMemcpyAsync(reinterpret_cast<char*>(d_sg) + qbytes, goals, qbytes, cudaMemcpyHostToDevice, s));

  // setStartGoalConfigs (RP:227-302): both roots must be valid, else the query fails immediately
  R_CUDA(ctx, launch_validity(ctx->tab_group, d_sg, NWB_F64, 2 * (int64_t)Q, d_valid, nullptr, nullptr, false, sp, s));
  std::vector<uint8_t> root_valid(2 * (size_t)Q);
  R_CUDA(ctx, cudaMemcpyAsync(root_valid.data(), d_valid, 2 * (size_t)Q, cudaMemcpyDeviceToHost, s));
  R_CUDA(ctx, cudaStreamSynchronize(s));
  int64_t launches = 1;
  std::vector<uint8_t> active(Q);
  std::vector<int32_t> roots(T), cnt(T);
  int32_t n_active = 0;
  for (int k = 0; k < Q; ++k) {
    std::memset(&results[k], 0, sizeof(nwb_plan_result));
    results[k].start_valid = root_valid[k];
    results[k].goal_valid = root_valid[Q + k];
    active[k] = root_valid[k] && root_valid[Q + k];
    n_active += active[k];
    roots[k] = 2 * k;                    // start tree of query k <- starts[k]
    roots[Q + k] = 2 * k + 1;            // goal tree            <- goals[k]
    cnt[k] = cnt[Q + k] = active[k] ? 1 : 0;
  }
  // insert the roots (index 0, predecessor 0: RP:236-258)
  R_CUDA(ctx, cudaMemcpyAsync(M + o_roots, roots.data(), sizeof(int32_t) * T, cudaMemcpyHostToDevice, s));
  R_CUDA(ctx, cudaMemcpyAsync(M + o_cnt, cnt.data(), sizeof(int32_t) * T, cudaMemcpyHostToDevice, s));
  R_CUDA(ctx, cudaMemcpyAsync(S.active, active.data(), Q, cudaMemcpyHostToDevice, s));
  R_CUDA(ctx, cudaMemcpyAsync(S.n_active, &n_active, sizeof(int32_t), cudaMemcpyHostToDevice, s));
  forest_append_kernel<<<T, 32, 0, s>>>(F, reinterpret_cast<int32_t*>(M + o_roots), d_sg, 1, reinterpret_cast<int32_t*>(M + o_cnt),
                                        reinterpret_cast<int32_t*>(M + o_zero), T, nullptr, S.overflow);
  if (P) {
    forest_root_hand_kernel<<<(T + 63) / 64, 64, 0, s>>>(F, T, P);
    ++launches;
  }
  R_CUDA(ctx, cudaGetLastError());
  ++launches;

  // ---- solveQuery main loop (RP:1067-1198) ---------------------------------------------------------------------------
  // Default: ONE launch, one warp per query, queries finish independently (rrt_persistent_kernel).
  // NWB_RRT_LOCKSTEP=1 selects the earlier driver: every iteration = 7-8 kernels over all still-active queries.
  static const bool lockstep = getenv("NWB_RRT_LOCKSTEP") != nullptr;
  if (!lockstep) {
    cudaError_t e = launch_rrt_persistent(pc, ctx->tab_group, sp, F, S, d_db, n_db, d_seeds, max_iter, Q, d_ma, P, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      ctx->err = std::string("nwb_solve_query_batch: ") + cudaGetErrorString(e);
      return NWB_E_CUDA;
    }
    ++launches;
  } else {
  if (!ctx->h_flag) R_CUDA(ctx, cudaMallocHost(&ctx->h_flag, 64));
  int32_t* h_flag = ctx->h_flag;                               // pinned: remaining active queries
  *h_flag = n_active;
  int check_every = 1, since_check = 0;
  for (int it = 0; it < max_iter && *h_flag > 0; ++it) {
    rrt_begin_kernel<<<Q, 32, 0, s>>>(d_db, n_db, d_seeds, it, pc.quirks, Q, S, d_qrand);
    // extendTree (RP:787-884)
    forest_nn_kernel<<<Q, 128, 0, s>>>(F, S.tree_a, S.active, d_qrand, Q, d_idx_a, d_near);
    cudaError_t e = launch_steer_project(pc, d_near, d_qrand, Q, d_qnew, d_status, nullptr, nullptr, nullptr, s);
    if (e == cudaSuccess && P) e = launch_ma_project(pc, d_ma, P, F, S, d_idx_a, d_near, d_qnew, d_status, Q, s);   // RP:829-837
    if (e == cudaSuccess) e = launch_validity(ctx->tab_group, d_qnew, NWB_F64, Q, d_valid, nullptr, nullptr, false, sp, s);
    rrt_extend_commit_kernel<<<Q, 32, 0, s>>>(F, S, d_status, d_valid, d_idx_a, d_qnew, Q, P);
    // connectTree (RP:890-1034) for the queries that extended, then the swap
    forest_nn_kernel<<<Q, 128, 0, s>>>(F, S.tree_b, S.walking, S.q_connect, Q, S.idx_b, S.cstart);
    if (e == cudaSuccess) e = launch_connect_forest(pc, ctx->tab_group, sp, F, S, Q, d_ma, P, s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
      ctx->err = std::string("nwb_solve_query_batch: ") + cudaGetErrorString(e);
      return NWB_E_CUDA;
    }
    launches += P ? 8 : 7;
    if (++since_check >= check_every || it + 1 == max_iter) {  // look at the active count now and then
      since_check = 0;
      check_every = std::min(16, check_every * 2);
      cudaMemcpyAsync(h_flag, S.n_active, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
      e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) {
          ctx->err = std::string("nwb_solve_query_batch: ") + cudaGetErrorString(e);
        return NWB_E_CUDA;
      }
    }
  }
  }

  int32_t overflow = 0;
  std::vector<int32_t> sizes(T), h_iter(Q), h_succ(Q), h_conn(Q), h_bridge(Q);
  R_CUDA(ctx, cudaMemcpyAsync(&overflow, S.overflow, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  R_CUDA(ctx, cudaMemcpyAsync(sizes.data(), F.size, sizeof(int32_t) * T, cudaMemcpyDeviceToHost, s));
  R_CUDA(ctx, cudaMemcpyAsync(h_iter.data(), S.iterations, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost, s));
  R_CUDA(ctx, cudaMemcpyAsync(h_succ.data(), S.success, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost, s));
  R_CUDA(ctx, cudaMemcpyAsync(h_conn.data(), S.connector, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost, s));
  R_CUDA(ctx, cudaMemcpyAsync(h_bridge.data(), S.bridge, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost, s));
  R_CUDA(ctx, cudaStreamSynchronize(s));

  // ---- writePath, path extraction part (RP:1303-1392): on the device, two synchronisations in total --------------
  for (int k = 0; k < Q; ++k) {
    nwb_plan_result& r = results[k];
    r.iterations = h_iter[k];
    r.success = h_succ[k];
    r.connector = r.success ? h_conn[k] : 0;
    r.bridge_other = r.success ? h_bridge[k] : 0;
    r.nodes_start = sizes[2 * k];
    r.nodes_goal = sizes[2 * k + 1];
    r.num_nodes_generated = r.success ? r.nodes_start + r.nodes_goal : 0;   // RP:1045: stays 0 unless a path is found
    r.raw_path_len = 0;
  }
  {
    int32_t* d_lens = reinterpret_cast<int32_t*>(M + o_cnt);                  // [Q][2] (the root-insertion counts are spent)
    R_CUDA(ctx, cudaMemsetAsync(d_lens, 0, sizeof(int32_t) * T, s));
    rrt_path_kernel<<<Q, 32, 0, s>>>(F, S.success, S.connector, S.bridge, Q, d_lens, nullptr, nullptr);
    std::vector<int32_t> lens(T);
    R_CUDA(ctx, cudaMemcpyAsync(lens.data(), d_lens, sizeof(int32_t) * T, cudaMemcpyDeviceToHost, s));
    R_CUDA(ctx, cudaStreamSynchronize(s));
    ++launches;
    std::vector<int64_t> offs(Q);
    int64_t total = 0;
    for (int k = 0; k < Q; ++k) {
      offs[k] = total;
      results[k].raw_path_len = lens[2 * k] + lens[2 * k + 1];
      total += results[k].raw_path_len;
    }
    if (raw_paths && total > 0) {
      const size_t need = sizeof(int64_t) * Q + 256 + sizeof(double) * NWB_NJ * (size_t)total;
      if (need > ctx->d_plan_bytes) {
        if (ctx->d_plan) cudaFree(ctx->d_plan);
        ctx->d_plan = nullptr;
        ctx->d_plan_bytes = 0;
        R_CUDA(ctx, cudaMalloc(&ctx->d_plan, std::max(need, (size_t)1 << 20)));
        ctx->d_plan_bytes = std::max(need, (size_t)1 << 20);
      }
      int64_t* d_offs = static_cast<int64_t*>(ctx->d_plan);
      double* d_rows = reinterpret_cast<double*>(static_cast<char*>(ctx->d_plan) + ((sizeof(int64_t) * Q + 255) & ~(size_t)255));
      R_CUDA(ctx, cudaMemcpyAsync(d_offs, offs.data(), sizeof(int64_t) * Q, cudaMemcpyHostToDevice, s));
      rrt_path_kernel<<<Q, 32, 0, s>>>(F, S.success, S.connector, S.bridge, Q, d_lens, d_offs, d_rows);
      std::vector<double> packed((size_t)total * NWB_NJ);
      R_CUDA(ctx, cudaMemcpyAsync(packed.data(), d_rows, sizeof(double) * NWB_NJ * (size_t)total, cudaMemcpyDeviceToHost, s));
      R_CUDA(ctx, cudaStreamSynchronize(s));
      ++launches;
      for (int k = 0; k < Q; ++k) {
        const int n = std::min<int>(results[k].raw_path_len, raw_path_capacity);
        if (n > 0)
          std::memcpy(raw_paths + (size_t)k * raw_path_capacity * NWB_NJ, packed.data() + (size_t)offs[k] * NWB_NJ, sizeof(double) * NWB_NJ * n);
      }
    }
  }
  ctx->last_launches = launches;
  if (overflow) {
    ctx->err = "nwb_solve_query_batch: a tree outgrew tree_capacity";
    return NWB_E_NOMEM;
  }
  return NWB_OK;
}

extern "C" {

int nwb_solve_query_batch(nwb_ctx* ctx, const double* starts, const double* goals, int64_t nq, const double* db, int32_t n_db,
                          const double* weights, const uint64_t* seeds, int32_t max_iter, double step, int32_t tree_capacity,
                          nwb_plan_result* results, double* raw_paths, int32_t raw_path_capacity) {
  return solve_impl(ctx, starts, goals, nq, db, n_db, weights, seeds, max_iter, step, tree_capacity, nullptr, 0, results, raw_paths,
                    raw_path_capacity);
}

int nwb_solve_query_constrained_batch(nwb_ctx* ctx, const double* starts, const double* goals, int64_t nq, const double* db,
                                      int32_t n_db, const double* weights, const uint64_t* seeds, int32_t max_iter, double step,
                                      int32_t tree_capacity, const double* hand_trajectories, int32_t n_points,
                                      nwb_plan_result* results, double* raw_paths, int32_t raw_path_capacity) {
  return solve_impl(ctx, starts, goals, nq, db, n_db, weights, seeds, max_iter, step, tree_capacity, hand_trajectories,
                    hand_trajectories ? n_points : 0, results, raw_paths, raw_path_capacity);
}

// path_shortcutter way-point selection (RP:1830-1951): every round evaluates ALL candidate edges of
// that round (current start -> raw[i], i = last .. 0) in one batched edge check, then takes the
// first valid one in the reference's scan order.  Returns the number of way-points (<= cap) or a
// negative code; -100 = the 500-round budget ran out (the reference then returns false).
int nwb_shortcut_path(nwb_ctx* ctx, const double* raw, int32_t n_raw, double* out, int32_t cap, int32_t* edges_checked) {
  if (!ctx) return NWB_E_INVALID;
  if (!raw || !out || n_raw < 1 || cap < 1) {
    ctx->err = "nwb_shortcut_path: bad argument";
    return NWB_E_INVALID;
  }
  const int W = n_raw, K = ctx->params.shortcut_waypoints;
  std::vector<double> qa((size_t)W * NWB_NJ), qb((size_t)W * NWB_NJ);
  std::vector<uint8_t> ok(W);
  auto row = [&](int i) { return raw + (size_t)i * NWB_NJ; };
  std::vector<const double*> sel;
  sel.push_back(row(0));
  const double* cur = row(0);
  int cur_idx = 0, budget = ctx->params.shortcut_max_rounds, edges = 0;
  bool reached = false;
  while (!reached && budget > 0) {
    // candidates in the reference's order: goal first (RP:1869-1876), then i = W-2 .. 0 (RP:1884-1916)
    std::vector<int> cand;
    cand.push_back(W - 1);
    for (int i = W - 2; i >= 0; --i) {
      if (i == cur_idx) break;
      cand.push_back(i);
    }
    const int n = (int)cand.size();
    for (int c = 0; c < n; ++c) {
      std::memcpy(&qa[(size_t)c * NWB_NJ], cur, sizeof(double) * NWB_NJ);
      std::memcpy(&qb[(size_t)c * NWB_NJ], row(cand[c]), sizeof(double) * NWB_NJ);
    }
    int rc = nwb_is_path_valid_batch(ctx, qa.data(), qb.data(), n, K, ok.data(), nullptr);
    if (rc != NWB_OK) return rc;
    int pick = -1;
    for (int c = 0; c < n; ++c) {
      ++edges;                                   // the serial reference stops at the first valid edge
      if (ok[c]) { pick = c; break; }
    }
    if (pick == 0) {
      sel.push_back(row(W - 1));
      reached = true;
    } else if (pick > 0) {
      cur_idx = cand[pick];
      cur = row(cur_idx);
      sel.push_back(cur);
    } else if (cur_idx + 1 < W) {                // stuck: advance one raw way-point (RP:1891-1897)
      cur_idx = cur_idx + 1;
      cur = row(cur_idx);
      sel.push_back(cur);
    }                                            // else: the reference's loop never meets i == current_start_index (the
    --budget;                                    // start already IS the last way-point): nothing is pushed, the budget runs down
  }
  if (edges_checked) *edges_checked = edges;
  if (budget == 0) return -100;
  const int n = (int)sel.size();
  for (int i = 0; i < std::min(n, (int)cap); ++i) std::memcpy(out + (size_t)i * NWB_NJ, sel[i], sizeof(double) * NWB_NJ);
  return n;
}

// interpolateShortPathWaypoints (RP:1956-1994): K points per segment; returns (W-1)(K+1)+1 rows.
int nwb_interpolate_path(const double* path, int32_t n_path, int32_t k, double* out, int32_t cap) {
  if (!path || !out || n_path < 1 || k < 0) return NWB_E_INVALID;
  int n = 0;
  auto push = [&](const double* r) {
    if (n < cap) std::memcpy(out + (size_t)n * NWB_NJ, r, sizeof(double) * NWB_NJ);
    ++n;
  };
  push(path);
  double sw[NWB_NJ], w[NWB_NJ];
  for (int i = 0; i + 1 < n_path; ++i) {
    const double* a = path + (size_t)i * NWB_NJ;
    const double* b = a + NWB_NJ;
    for (int j = 0; j < NWB_NJ; ++j) sw[j] = (b[j] - a[j]) / (k + 1);
    for (int m = 1; m <= k + 1; ++m) {
      for (int j = 0; j < NWB_NJ; ++j) w[j] = a[j] + sw[j] * m;
      push(w);
    }
  }
  return n;
}

}  // extern "C"
