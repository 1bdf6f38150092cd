// rrt_driver.cu - host RRT-Connect driver over the batched device operators.
//
// Replaces RRT_Planner::setStartGoalConfigs / solveQuery / extendTree / connectTree /
// addConfigtoTree / writePath (rrt_connect_planner/src/rrt_planner.cpp:227-302, 1039-1212, 787-884,
// 890-1034, 749-783, 1303-1392) and path_shortcutter / interpolateShortPathWaypoints (RP:1830-1994)
// for nq INDEPENDENT planning queries advanced in lock step: iteration i of every still-active
// query is one pass of
//     random DB row -> nearest neighbour -> steer -> double-support projection -> validity -> append
//     -> nearest neighbour in the other tree -> connect walk -> append -> swap
// where each arrow is one kernel over all queries.  The host only keeps the control flow (which
// queries are active, which tree is which); every distance, IK step and validity test runs on the
// device.  A single query is the nq = 1 case (latency bound: RRT-Connect is serial by construction,
// SURVEY.md section 8e "a single RRT query: replicas only").
//
// Trees live in one pooled forest (2 trees per query, fixed capacity each):
//   soa float [T][21][C] | aos double [T][C][21] | pred int32 [T][C] | size int32 [T]
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "nwb_internal.h"
#include "planner_core.cuh"

namespace nwb {
void make_plan_const(const nwb_ctx* ctx, const double* weights, double step, PlanConst& pc);

struct ForestView {
  float* soa;
  double* aos;
  int32_t* pred;
  int32_t* size;
  int64_t cap;
};

__device__ __forceinline__ bool lex_less2(double da, int ia, double db, int ib) {
  return (da < db) || (da == db && (unsigned)ia < (unsigned)ib);
}

// q_rand[k] = DB row floor(U * (n-1)) (quirk C3) or floor(U * n); U = Philox(seed_k; iteration)
__global__ void rrt_random_rows_kernel(const double* __restrict__ db, int n_db, const uint64_t* __restrict__ seeds, int iteration,
                                       unsigned quirks, int nq, double* __restrict__ q_rand) {
  int k = blockIdx.x;
  if (k >= nq) return;
  __shared__ int row;
  if (threadIdx.x == 0) {
    uint32_t r[4];
    uint64_t s = seeds[k];
    philox4x32_10((uint32_t)iteration, 0u, 0u, kStreamPlanner, (uint32_t)s, (uint32_t)(s >> 32), r);
    double u = u01(r[0]);
    row = (quirks & NWB_QUIRK_DB_INDEX_FLOOR) ? (int)floor(0.0 + (double)(n_db - 1 - 0) * u) : (int)floor((double)n_db * u);
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

// after steer/project + validity: decide which queries extend (RP:846-870) and set up the connect edges
__global__ void rrt_extend_commit_kernel(const int32_t* __restrict__ status, const uint8_t* __restrict__ valid,
                                         const uint8_t* __restrict__ active, int nq, int32_t* __restrict__ ext_state,
                                         int32_t* __restrict__ ext_count) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nq) return;
  int st = NWB_TRAPPED;
  if (active[k] && status[k] != NWB_TRAPPED && valid[k]) st = status[k];
  ext_state[k] = st;
  ext_count[k] = (st != NWB_TRAPPED) ? 1 : 0;
}

}  // namespace nwb

using namespace nwb;

namespace {
#define R_CUDA(ctx, call)                                                      \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) {                                                   \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);         \
      return NWB_E_CUDA;                                                       \
    }                                                                          \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
  template <class T> T* as() const { return static_cast<T*>(p); }
};
}  // namespace

extern "C" {

int nwb_solve_query_batch(nwb_ctx* ctx, const double* starts, const double* goals, int64_t nq, const double* db, int32_t n_db,
                          const double* weights, const uint64_t* seeds, int32_t max_iter, double step, int32_t tree_capacity,
                          nwb_plan_result* results, double* raw_paths, int32_t raw_path_capacity) {
  if (!ctx) return NWB_E_INVALID;
  if (nq < 0 || (nq > 0 && (!starts || !goals || !db || !seeds || !results)) || n_db < 2 || max_iter < 0 || !(step > 0) ||
      (raw_paths && raw_path_capacity <= 0)) {
    ctx->err = "nwb_solve_query_batch: bad argument";
    return NWB_E_INVALID;
  }
  if (nq == 0) return NWB_OK;
  R_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int Q = (int)nq, T = 2 * Q;
  const int64_t C = tree_capacity > 0 ? tree_capacity : 4096;
  const int max_steps = 64;                                // connect nodes per launch; longer walks continue in another round
  const bool sp = ctx->params.srdf_trim != 0 && !ctx->force_table_pairs;
  PlanConst pc;
  make_plan_const(ctx, weights, step, pc);

  DevBuf b_soa, b_aos, b_pred, b_size, b_db, b_seeds, b_qrand, b_near, b_qnew, b_status, b_valid, b_active, b_tree_a, b_tree_b,
      b_idx_a, b_idx_b, b_ext_state, b_ext_count, b_first, b_overflow, b_chain, b_nadd, b_cstat, b_cstart, b_pfirst, b_sg;
  const size_t qbytes = sizeof(double) * NWB_NJ * Q;
  R_CUDA(ctx, b_soa.alloc(sizeof(float) * NWB_NJ * C * T));
  R_CUDA(ctx, b_aos.alloc(sizeof(double) * NWB_NJ * C * T));
  R_CUDA(ctx, b_pred.alloc(sizeof(int32_t) * C * T));
  R_CUDA(ctx, b_size.alloc(sizeof(int32_t) * T));
  R_CUDA(ctx, b_db.alloc(sizeof(double) * NWB_NJ * n_db));
  R_CUDA(ctx, b_seeds.alloc(sizeof(uint64_t) * Q));
  R_CUDA(ctx, b_qrand.alloc(qbytes));
  R_CUDA(ctx, b_near.alloc(qbytes));
  R_CUDA(ctx, b_qnew.alloc(qbytes));
  R_CUDA(ctx, b_cstart.alloc(qbytes));
  R_CUDA(ctx, b_status.alloc(sizeof(int32_t) * Q));
  R_CUDA(ctx, b_valid.alloc(2 * (size_t)Q));
  R_CUDA(ctx, b_active.alloc(Q));
  R_CUDA(ctx, b_tree_a.alloc(sizeof(int32_t) * Q));
  R_CUDA(ctx, b_tree_b.alloc(sizeof(int32_t) * Q));
  R_CUDA(ctx, b_idx_a.alloc(sizeof(int32_t) * Q));
  R_CUDA(ctx, b_idx_b.alloc(sizeof(int32_t) * Q));
  R_CUDA(ctx, b_ext_state.alloc(sizeof(int32_t) * Q));
  R_CUDA(ctx, b_ext_count.alloc(sizeof(int32_t) * Q));
  R_CUDA(ctx, b_first.alloc(sizeof(int32_t) * Q));
  R_CUDA(ctx, b_overflow.alloc(sizeof(int32_t)));
  R_CUDA(ctx, b_chain.alloc(qbytes * max_steps));
  R_CUDA(ctx, b_nadd.alloc(sizeof(int32_t) * Q));
  R_CUDA(ctx, b_cstat.alloc(sizeof(int32_t) * Q));
  R_CUDA(ctx, b_pfirst.alloc(sizeof(int32_t) * Q));
  R_CUDA(ctx, b_sg.alloc(2 * qbytes));
  ForestView F{b_soa.as<float>(), b_aos.as<double>(), b_pred.as<int32_t>(), b_size.as<int32_t>(), C};

  R_CUDA(ctx, cudaMemsetAsync(b_size.p, 0, sizeof(int32_t) * T, s));
  R_CUDA(ctx, cudaMemsetAsync(b_near.p, 0, qbytes, s));       // inactive slots still flow through the kernels: keep them finite
  R_CUDA(ctx, cudaMemsetAsync(b_qnew.p, 0, qbytes, s));
  R_CUDA(ctx, cudaMemsetAsync(b_cstart.p, 0, qbytes, s));
  R_CUDA(ctx, cudaMemsetAsync(b_qrand.p, 0, qbytes, s));
  R_CUDA(ctx, cudaMemsetAsync(b_overflow.p, 0, sizeof(int32_t), s));
  R_CUDA(ctx, cudaMemcpyAsync(b_db.p, db, sizeof(double) * NWB_NJ * n_db, cudaMemcpyHostToDevice, s));
  R_CUDA(ctx, cudaMemcpyAsync(b_seeds.p, seeds, sizeof(uint64_t) * Q, cudaMemcpyHostToDevice, s));
  R_CUDA(ctx, cudaMemcpyAsync(b_sg.p, starts, qbytes, cudaMemcpyHostToDevice, s));
  R_CUDA(ctx, cudaMemcpyAsync(b_sg.as<char>() + qbytes, goals, qbytes, cudaMemcpyHostToDevice, s));

  // setStartGoalConfigs (RP:227-302): both roots must be valid, else the query fails immediately
  R_CUDA(ctx, launch_validity(ctx->tab_group, b_sg.p, NWB_F64, 2 * (int64_t)Q, b_valid.as<uint8_t>(), nullptr, nullptr, false, sp, s));
  std::vector<uint8_t> root_valid(2 * (size_t)Q);
  R_CUDA(ctx, cudaMemcpyAsync(root_valid.data(), b_valid.p, 2 * (size_t)Q, cudaMemcpyDeviceToHost, s));
  R_CUDA(ctx, cudaStreamSynchronize(s));

  std::vector<uint8_t> active(Q);
  std::vector<int32_t> tree_a(Q), tree_b(Q), ones(Q, 1), minus1(Q, -1), roots(T);
  std::vector<int> swap_flag(Q, 0);
  int64_t launches = 1;
  for (int k = 0; k < Q; ++k) {
    std::memset(&results[k], 0, sizeof(nwb_plan_result));
    results[k].start_valid = root_valid[k];
    results[k].goal_valid = root_valid[Q + k];
    active[k] = root_valid[k] && root_valid[Q + k];
    roots[k] = 2 * k;                    // start tree of query k
    roots[Q + k] = 2 * k + 1;            // goal tree
  }
  {  // insert the roots: tree 2k <- start_k, tree 2k+1 <- goal_k (index 0, predecessor 0: RP:236-258)
    DevBuf b_roots, b_cnt, b_zero;
    R_CUDA(ctx, b_roots.alloc(sizeof(int32_t) * T));
    R_CUDA(ctx, b_cnt.alloc(sizeof(int32_t) * T));
    R_CUDA(ctx, b_zero.alloc(sizeof(int32_t) * T));
    std::vector<int32_t> cnt(T), zero(T, 0);
    for (int k = 0; k < Q; ++k) cnt[k] = cnt[Q + k] = active[k] ? 1 : 0;
    R_CUDA(ctx, cudaMemcpyAsync(b_roots.p, roots.data(), sizeof(int32_t) * T, cudaMemcpyHostToDevice, s));
    R_CUDA(ctx, cudaMemcpyAsync(b_cnt.p, cnt.data(), sizeof(int32_t) * T, cudaMemcpyHostToDevice, s));
    R_CUDA(ctx, cudaMemcpyAsync(b_zero.p, zero.data(), sizeof(int32_t) * T, cudaMemcpyHostToDevice, s));
    forest_append_kernel<<<T, 32, 0, s>>>(F, b_roots.as<int32_t>(), b_sg.as<double>(), 1, b_cnt.as<int32_t>(), b_zero.as<int32_t>(), T,
                                          nullptr, b_overflow.as<int32_t>());
    R_CUDA(ctx, cudaGetLastError());
    R_CUDA(ctx, cudaStreamSynchronize(s));
    ++launches;
  }

  std::vector<int32_t> h_ext(Q), h_cstat(Q), h_nadd(Q), h_idx_b(Q), h_first(Q), h_pfirst(Q), walk_added(Q), walk_near(Q);
  std::vector<uint8_t> walking(Q);
  int n_active = 0;
  for (int k = 0; k < Q; ++k) n_active += active[k];

  for (int it = 0; it < max_iter && n_active > 0; ++it) {
    for (int k = 0; k < Q; ++k) {
      tree_a[k] = 2 * k + swap_flag[k];
      tree_b[k] = 2 * k + 1 - swap_flag[k];
      if (active[k]) results[k].iterations = it + 1;
    }
    R_CUDA(ctx, cudaMemcpyAsync(b_active.p, active.data(), Q, cudaMemcpyHostToDevice, s));
    R_CUDA(ctx, cudaMemcpyAsync(b_tree_a.p, tree_a.data(), sizeof(int32_t) * Q, cudaMemcpyHostToDevice, s));
    R_CUDA(ctx, cudaMemcpyAsync(b_tree_b.p, tree_b.data(), sizeof(int32_t) * Q, cudaMemcpyHostToDevice, s));
    // ---- extendTree (RP:787-884) -----------------------------------------------------------------
    rrt_random_rows_kernel<<<Q, 32, 0, s>>>(b_db.as<double>(), n_db, b_seeds.as<uint64_t>(), it, pc.quirks, Q, b_qrand.as<double>());
    forest_nn_kernel<<<Q, 128, 0, s>>>(F, b_tree_a.as<int32_t>(), b_active.as<uint8_t>(), b_qrand.as<double>(), Q, b_idx_a.as<int32_t>(),
                                       b_near.as<double>());
    R_CUDA(ctx, launch_steer_project(pc, b_near.as<double>(), b_qrand.as<double>(), Q, b_qnew.as<double>(), b_status.as<int32_t>(),
                                     nullptr, nullptr, s));
    R_CUDA(ctx, launch_validity(ctx->tab_group, b_qnew.p, NWB_F64, Q, b_valid.as<uint8_t>(), nullptr, nullptr, false, sp, s));
    rrt_extend_commit_kernel<<<(Q + 127) / 128, 128, 0, s>>>(b_status.as<int32_t>(), b_valid.as<uint8_t>(), b_active.as<uint8_t>(), Q,
                                                             b_ext_state.as<int32_t>(), b_ext_count.as<int32_t>());
    forest_append_kernel<<<Q, 32, 0, s>>>(F, b_tree_a.as<int32_t>(), b_qnew.as<double>(), 1, b_ext_count.as<int32_t>(),
                                          b_idx_a.as<int32_t>(), Q, b_first.as<int32_t>(), b_overflow.as<int32_t>());
    launches += 6;
    R_CUDA(ctx, cudaMemcpyAsync(h_ext.data(), b_ext_state.p, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost, s));
    R_CUDA(ctx, cudaStreamSynchronize(s));
    // ---- connectTree (RP:890-1034) for the queries that extended ---------------------------------------
    int n_walk = 0;
    for (int k = 0; k < Q; ++k) {
      walking[k] = active[k] && h_ext[k] != NWB_TRAPPED;
      n_walk += walking[k];
      walk_added[k] = 0;
    }
    if (n_walk > 0) {
      R_CUDA(ctx, cudaMemcpyAsync(b_active.p, walking.data(), Q, cudaMemcpyHostToDevice, s));
      // q_connect = the node just appended to Ta = q_new; nearest neighbour in Tb is the walk's first node
      forest_nn_kernel<<<Q, 128, 0, s>>>(F, b_tree_b.as<int32_t>(), b_active.as<uint8_t>(), b_qnew.as<double>(), Q, b_idx_b.as<int32_t>(),
                                         b_cstart.as<double>());
      R_CUDA(ctx, cudaMemcpyAsync(h_idx_b.data(), b_idx_b.p, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost, s));
      R_CUDA(ctx, cudaStreamSynchronize(s));
      ++launches;
      for (int k = 0; k < Q; ++k) {
        h_pfirst[k] = h_idx_b[k];
        walk_near[k] = h_idx_b[k];
      }
      while (n_walk > 0) {
        // only the still-walking edges take steps: mask the others by making them zero-length REACHED... simpler:
        // run every edge slot; results of non-walking slots are ignored and never appended (count forced to 0)
        R_CUDA(ctx, launch_connect(pc, ctx->tab_group, sp, b_cstart.as<double>(), b_qnew.as<double>(), Q, max_steps, b_chain.as<double>(),
                                   b_nadd.as<int32_t>(), b_cstat.as<int32_t>(), s));
        R_CUDA(ctx, cudaMemcpyAsync(h_nadd.data(), b_nadd.p, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost, s));
        R_CUDA(ctx, cudaMemcpyAsync(h_cstat.data(), b_cstat.p, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost, s));
        R_CUDA(ctx, cudaStreamSynchronize(s));
        for (int k = 0; k < Q; ++k)
          if (!walking[k]) h_nadd[k] = 0;
        R_CUDA(ctx, cudaMemcpyAsync(b_nadd.p, h_nadd.data(), sizeof(int32_t) * Q, cudaMemcpyHostToDevice, s));
        R_CUDA(ctx, cudaMemcpyAsync(b_pfirst.p, h_pfirst.data(), sizeof(int32_t) * Q, cudaMemcpyHostToDevice, s));
        forest_append_kernel<<<Q, 32, 0, s>>>(F, b_tree_b.as<int32_t>(), b_chain.as<double>(), max_steps, b_nadd.as<int32_t>(),
                                              b_pfirst.as<int32_t>(), Q, b_first.as<int32_t>(), b_overflow.as<int32_t>());
        R_CUDA(ctx, cudaMemcpyAsync(h_first.data(), b_first.p, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost, s));
        R_CUDA(ctx, cudaStreamSynchronize(s));
        launches += 2;
        n_walk = 0;
        for (int k = 0; k < Q; ++k) {
          if (!walking[k]) continue;
          const int n = h_nadd[k];
          // q_near of writePath (RP:1000-1001): the node the REACHED step started from
          if (n >= 2) walk_near[k] = h_first[k] + n - 2;
          else if (n == 1 && walk_added[k] > 0) walk_near[k] = h_pfirst[k];
          walk_added[k] += n;
          if (h_cstat[k] == -1) {          // budget exhausted: continue from the last appended node
            h_pfirst[k] = h_first[k] + n - 1;
            R_CUDA(ctx, cudaMemcpyAsync(b_cstart.as<double>() + (size_t)k * NWB_NJ,
                                        b_chain.as<double>() + ((size_t)k * max_steps + n - 1) * NWB_NJ, sizeof(double) * NWB_NJ,
                                        cudaMemcpyDeviceToDevice, s));
            ++n_walk;
          } else {
            walking[k] = 0;
            if (h_cstat[k] == NWB_REACHED) {
              results[k].success = 1;
              results[k].connector = swap_flag[k] ? 2 : 1;      // 1: goal tree connected to start tree (RP:1101)
              results[k].bridge_other = walk_near[k];
              active[k] = 0;
              --n_active;
            }
          }
        }
      }
    }
    for (int k = 0; k < Q; ++k)
      if (active[k]) swap_flag[k] ^= 1;         // trees swap every iteration regardless of outcome (RP:1125-1192)
  }

  int32_t overflow = 0;
  std::vector<int32_t> sizes(T);
  R_CUDA(ctx, cudaMemcpyAsync(&overflow, b_overflow.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  R_CUDA(ctx, cudaMemcpyAsync(sizes.data(), b_size.p, sizeof(int32_t) * T, cudaMemcpyDeviceToHost, s));
  R_CUDA(ctx, cudaStreamSynchronize(s));

  // ---- writePath, path extraction part (RP:1303-1392) -------------------------------------------------------
  std::vector<double> cfg;
  std::vector<int32_t> pred;
  for (int k = 0; k < Q; ++k) {
    nwb_plan_result& r = results[k];
    r.nodes_start = sizes[2 * k];
    r.nodes_goal = sizes[2 * k + 1];
    r.num_nodes_generated = r.success ? r.nodes_start + r.nodes_goal : 0;   // RP:1045: stays 0 unless a path is found
    r.raw_path_len = 0;
    if (!r.success) continue;
    int idx_s, idx_g;
    if (r.connector == 1) { idx_s = r.nodes_start - 1; idx_g = r.bridge_other; }
    else { idx_s = r.bridge_other; idx_g = r.nodes_goal - 1; }
    std::vector<std::vector<double>> path;
    for (int side = 0; side < 2; ++side) {
      const int t = 2 * k + side, n = sizes[t];
      cfg.resize((size_t)n * NWB_NJ);
      pred.resize(n);
      R_CUDA(ctx, cudaMemcpyAsync(cfg.data(), F.aos + (size_t)t * C * NWB_NJ, sizeof(double) * NWB_NJ * n, cudaMemcpyDeviceToHost, s));
      R_CUDA(ctx, cudaMemcpyAsync(pred.data(), F.pred + (size_t)t * C, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
      R_CUDA(ctx, cudaStreamSynchronize(s));
      std::vector<std::vector<double>> branch;
      int i = side == 0 ? idx_s : idx_g;
      while (i > 0) {
        branch.emplace_back(cfg.begin() + (size_t)i * NWB_NJ, cfg.begin() + (size_t)(i + 1) * NWB_NJ);
        i = pred[i];
      }
      branch.emplace_back(cfg.begin(), cfg.begin() + NWB_NJ);
      if (side == 0) path.assign(branch.rbegin(), branch.rend());        // root ... bridge
      else path.insert(path.end(), branch.begin(), branch.end());        // bridge ... goal root
    }
    r.raw_path_len = (int32_t)path.size();
    if (raw_paths)
      for (int i = 0; i < std::min<int>(r.raw_path_len, raw_path_capacity); ++i)
        std::memcpy(raw_paths + ((size_t)k * raw_path_capacity + i) * NWB_NJ, path[i].data(), sizeof(double) * NWB_NJ);
  }
  ctx->last_launches = launches;
  if (overflow) {
    ctx->err = "nwb_solve_query_batch: a tree outgrew tree_capacity";
    return NWB_E_NOMEM;
  }
  return NWB_OK;
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
    } else {                                     // stuck: advance one raw way-point (RP:1891-1897)
      cur_idx = cur_idx + 1;
      cur = row(cur_idx);
      sel.push_back(cur);
    }
    --budget;
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
