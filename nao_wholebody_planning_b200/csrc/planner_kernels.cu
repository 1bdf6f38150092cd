// planner_kernels.cu - batched local-planner operators for sm_100a:
//   steer_project   generate_q_new -> enforce_DS_Constraints         (RP:424-494, 498-593)
//   connect         the connectTree edge walk, one thread per edge   (RP:926-1031)
//   edge_check      interpolateWaypoints + isPathValid               (RP:1677-1774, 1876, 1905)
//   sample          sample_stable_configs, database branch           (SCG:278-480)
// RP = rrt_connect_planner/src/rrt_planner.cpp, SCG = .../stable_config_generator.cpp.
// The per-thread building blocks live in planner_core.cuh (fp64 IK) and validity_core.cuh (fp32
// validity); these kernels only lay the work out: one unit (edge step chain, way-point, sample) per
// thread, coalesced output, no inter-thread communication except the per-edge "first bad way-point"
// minimum and the compaction counter of the sampler.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdlib>

#include "arm_core.cuh"
#include "planner_core.cuh"
#include "validity_core.cuh"

namespace nwb {

// Point table of one configuration in thread-private storage: registers when every slot is a
// compile-time constant (compiled-in pair list), local memory under the table-driven pair list.
struct LocalStore {
  float v[nwbm::kPointFloats];
  __device__ __forceinline__ void put(int slot, float x) { v[slot] = x; }
  __device__ __forceinline__ float get(int slot) const { return v[slot]; }
};

template <bool COLLISION_ONLY, bool STATIC>
__device__ __forceinline__ ValidityResult<float> eval_one(const ValidityTables& tab, const double* q) {
  float qf[NWB_NJ];
#pragma unroll
  for (int j = 0; j < NWB_NJ; ++j) qf[j] = (float)q[j];
  LocalStore st;
  constexpr int P = STATIC ? (COLLISION_ONLY ? PAIRS_STATIC_ALL : PAIRS_STATIC_GROUP) : PAIRS_TABLE;
  // Two instantiations behind a uniform branch: the hierarchy walk of large scenes raises the register pressure (robot
  // bounding box, a second copy of the leaf test) enough to push the point table of the compiled-in pair list out of
  // registers (edge check: 560 B of stack against 8 B), and the scenes the reference builds never need it.
  if (tab.scene.n_nodes > 0)
    return evaluate_config<float, LocalStore, false, COLLISION_ONLY, P, /*ENV_CULL=*/true, /*BIG_SCENE=*/true>(tab, qf, st, nullptr, nullptr);
  return evaluate_config<float, LocalStore, false, COLLISION_ONLY, P, /*ENV_CULL=*/true, /*BIG_SCENE=*/false>(tab, qf, st, nullptr, nullptr);
}

// ---- steer + project ----------------------------------------------------------------------------
#ifndef NWB_STEER_MIN_BLOCKS
#define NWB_STEER_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(128, NWB_STEER_MIN_BLOCKS)
steer_project_kernel(const __grid_constant__ PlanConst pc, const double* __restrict__ q_near, const double* __restrict__ q_target,
                     int64_t n, double* __restrict__ q_out, int32_t* __restrict__ status, int32_t* __restrict__ ds_status,
                     int32_t* __restrict__ iters) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[NWB_NJ], b[NWB_NJ], q[NWB_NJ];
  for (int j = 0; j < NWB_NJ; ++j) {
    a[j] = q_near[i * NWB_NJ + j];
    b[j] = q_target[i * NWB_NJ + j];
  }
  int st = generate_q_new(a, b, pc, q);
  int it = 0;
  int ds = enforce_ds_constraints(q, pc, &it);
  if (ds_status) ds_status[i] = ds;
  if (iters) iters[i] = it;
  if (ds == NWB_TRAPPED) {
    status[i] = NWB_TRAPPED;
    return;
  }
  if (ds == NWB_ADVANCED) st = NWB_ADVANCED;
  status[i] = st;
  for (int j = 0; j < NWB_NJ; ++j) q_out[i * NWB_NJ + j] = q[j];
}

// The same operator as two launches (large batches): the first classifies every pair - joint limits, already in double
// support, or a foot target the solve can never converge to (left_leg_target_feasible) - and packs the pairs that need
// the Newton-Raphson solve into a dense list; the second runs one solve per thread on full warps.  generate_q_new is
// recomputed in the second launch (cheap, and rows of TRAPPED pairs must stay untouched in q_out).
__device__ __forceinline__ unsigned steer_warp_append(bool keep, unsigned* counter) {
  const unsigned mask = __ballot_sync(0xffffffffu, keep);
  if (!mask) return 0;
  const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
}
__global__ void __launch_bounds__(128)
steer_filter_kernel(const __grid_constant__ PlanConst pc, const double* __restrict__ q_near, const double* __restrict__ q_target,
                    int64_t n, double* __restrict__ q_out, int32_t* __restrict__ status, int32_t* __restrict__ ds_status,
                    int32_t* __restrict__ iters, uint32_t* __restrict__ work, unsigned* __restrict__ counter) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool need = false;
  if (i < n) {
    double a[NWB_NJ], b[NWB_NJ], q[NWB_NJ];
    for (int j = 0; j < NWB_NJ; ++j) {
      a[j] = q_near[i * NWB_NJ + j];
      b[j] = q_target[i * NWB_NJ + j];
    }
    const int st = generate_q_new(a, b, pc, q);
    int ds = NWB_TRAPPED, it = 0;
    if (limits_ok(q, pc.jmin, pc.jmax, NWB_NJ)) {
      double rl[5];
#pragma unroll
      for (int m = 0; m < 5; ++m) rl[m] = -q[m];
      DFrame hip;
      right_leg_fk(rl, hip);
      DFrame foot = hip;
      left_leg_fk(q + 16, foot, nullptr, nullptr);
      const double des[3] = {nwbm::kLeftSoleTarget[0], nwbm::kLeftSoleTarget[1], nwbm::kLeftSoleTarget[2]};
      const double act[3] = {foot.p.x, foot.p.y, foot.p.z};
      double max_err = 0.0;
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        double e = des[m] - act[m];
        if (!(pc.quirks & NWB_QUIRK_SIGNED_FOOT_ERROR)) e = fabs(e);
        if (e > max_err) max_err = e;
      }
      if (!(max_err > pc.ds_tolerance)) {
        ds = NWB_REACHED;
      } else {
        DFrame target;
        desired_left_leg_pose(hip, target);
        if (left_leg_target_feasible(target, pc.ik_eps)) need = true;
        else it = pc.ik_maxiter;                            // what the full loop would report
      }
    }
    if (!need) {
      if (ds_status) ds_status[i] = ds;
      if (iters) iters[i] = it;
      status[i] = ds == NWB_TRAPPED ? NWB_TRAPPED : st;
      if (ds != NWB_TRAPPED)
        for (int j = 0; j < NWB_NJ; ++j) q_out[i * NWB_NJ + j] = q[j];
    }
  }
  const unsigned slot = steer_warp_append(need, counter);
  if (need) work[slot] = (uint32_t)i;
}
// The solve is split once more, as in the database sampler: the head runs the first kSteerHead Newton-Raphson iterations
// of every packed pair; a solve that is still running hands its state (5 joints + the 5x5 right singular vectors,
// struct-of-arrays) to a second packed list and the tail kernel resumes it, again on full warps.  After the exact
// rejection rule a quarter of the packed solves still fail the slow way (all `maxiter` iterations): without the split
// every warp of 32 solves contains one and runs 100 iterations.  counters[0] = |work|, counters[1] = |tail|.
constexpr int kSteerHead = 10;
constexpr int kSteerTailState = 30;
__device__ __forceinline__ void steer_finish(const PlanConst& pc, int64_t i, int it, double* q, const double* ll,
                                             double* __restrict__ q_out, int32_t* __restrict__ status,
                                             int32_t* __restrict__ ds_status, int32_t* __restrict__ iters) {
  const bool ok = it >= 0 && limits_ok(ll, pc.jmin + 16, pc.jmax + 16, 5);
  if (iters) iters[i] = it >= 0 ? it + 1 : pc.ik_maxiter;
  if (ds_status) ds_status[i] = ok ? NWB_ADVANCED : NWB_TRAPPED;
  status[i] = ok ? NWB_ADVANCED : NWB_TRAPPED;
  if (ok) {
#pragma unroll
    for (int m = 0; m < 5; ++m) q[16 + m] = ll[m];
    for (int j = 0; j < NWB_NJ; ++j) q_out[i * NWB_NJ + j] = q[j];
  }
}
__global__ void __launch_bounds__(64, 2 * NWB_STEER_MIN_BLOCKS)   // 12 warps per SM, as the one-launch form
steer_solve_kernel(const __grid_constant__ PlanConst pc, const double* __restrict__ q_near, const double* __restrict__ q_target,
                   const uint32_t* __restrict__ work, unsigned* __restrict__ counters, uint32_t* __restrict__ tail_i,
                   double* __restrict__ tail_state, unsigned tail_cap, double* __restrict__ q_out,
                   int32_t* __restrict__ status, int32_t* __restrict__ ds_status, int32_t* __restrict__ iters) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = w < (int64_t)counters[0];
  if (__ballot_sync(0xffffffffu, active) == 0u) return;
  int64_t i = 0;
  double q[NWB_NJ], ll[5], V[5][5];
  DFrame target;
  int it = -1;
  const int head = pc.ik_maxiter < kSteerHead ? pc.ik_maxiter : kSteerHead;
  if (active) {
    i = work[w];
    double a[NWB_NJ], b[NWB_NJ];
    for (int j = 0; j < NWB_NJ; ++j) {
      a[j] = q_near[i * NWB_NJ + j];
      b[j] = q_target[i * NWB_NJ + j];
    }
    generate_q_new(a, b, pc, q);
    double rl[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      rl[m] = -q[m];
      ll[m] = q[16 + m];
    }
    DFrame hip;
    right_leg_fk(rl, hip);
    desired_left_leg_pose(hip, target);
    nr_identity(V);
    it = left_leg_ik_nr_span(target, ll, V, 0, head, pc.ik_eps, pc.pinv_eps);
  }
  const bool more = active && it < 0 && head < pc.ik_maxiter;
  const unsigned slot = steer_warp_append(more, counters + 1);
  if (more) {
    if (slot < tail_cap) {
      tail_i[slot] = (uint32_t)i;
#pragma unroll
      for (int m = 0; m < 5; ++m) tail_state[(size_t)m * tail_cap + slot] = ll[m];
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) tail_state[(size_t)(5 + 5 * a + b) * tail_cap + slot] = V[a][b];
      return;
    }
    it = left_leg_ik_nr_span(target, ll, V, head, pc.ik_maxiter, pc.ik_eps, pc.pinv_eps);   // list full: finish in place
  }
  if (active) steer_finish(pc, i, it, q, ll, q_out, status, ds_status, iters);
}
__global__ void __launch_bounds__(64, 2 * NWB_STEER_MIN_BLOCKS)
steer_tail_kernel(const __grid_constant__ PlanConst pc, const double* __restrict__ q_near, const double* __restrict__ q_target,
                  const unsigned* __restrict__ counters, const uint32_t* __restrict__ tail_i, const double* __restrict__ tail_state,
                  unsigned tail_cap, double* __restrict__ q_out, int32_t* __restrict__ status, int32_t* __restrict__ ds_status,
                  int32_t* __restrict__ iters) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned queued = counters[1] < tail_cap ? counters[1] : tail_cap;
  if (w >= (int64_t)queued) return;
  const int64_t i = tail_i[w];
  double a[NWB_NJ], b[NWB_NJ], q[NWB_NJ], rl[5], ll[5], V[5][5];
  for (int j = 0; j < NWB_NJ; ++j) {
    a[j] = q_near[i * NWB_NJ + j];
    b[j] = q_target[i * NWB_NJ + j];
  }
  generate_q_new(a, b, pc, q);
#pragma unroll
  for (int m = 0; m < 5; ++m) rl[m] = -q[m];
  DFrame hip, target;
  right_leg_fk(rl, hip);
  desired_left_leg_pose(hip, target);
#pragma unroll
  for (int m = 0; m < 5; ++m) ll[m] = tail_state[(size_t)m * tail_cap + w];
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c) V[r][c] = tail_state[(size_t)(5 + 5 * r + c) * tail_cap + w];
  const int it = left_leg_ik_nr_span(target, ll, V, kSteerHead, pc.ik_maxiter, pc.ik_eps, pc.pinv_eps);
  steer_finish(pc, i, it, q, ll, q_out, status, ds_status, iters);
}

// ---- connect: one thread walks one edge -------------------------------------------------------------
// From q_start repeat {generate_q_new -> enforce_DS -> isConfigValid -> append} until REACHED or
// TRAPPED; step k+1 starts from the PROJECTED node of step k (RP:1004), so the walk is serial by
// construction and the parallelism is across edges.
template <bool STATIC>
__global__ void __launch_bounds__(64)
connect_kernel(const __grid_constant__ PlanConst pc, const __grid_constant__ ValidityTables tab,
               const double* __restrict__ q_start, const double* __restrict__ q_connect, int64_t n_edges, int max_steps,
               double* __restrict__ nodes_out /*[n_edges][max_steps][21]*/, int32_t* __restrict__ n_added,
               int32_t* __restrict__ status) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  double cur[NWB_NJ], tgt[NWB_NJ], q[NWB_NJ];
  for (int j = 0; j < NWB_NJ; ++j) {
    cur[j] = q_start[e * NWB_NJ + j];
    tgt[j] = q_connect[e * NWB_NJ + j];
  }
  int state = NWB_ADVANCED, added = 0;
  while (state == NWB_ADVANCED && added < max_steps) {
    state = generate_q_new(cur, tgt, pc, q);
    int ds = enforce_ds_constraints(q, pc, nullptr);
    if (ds == NWB_TRAPPED) { state = NWB_TRAPPED; break; }
    if (ds == NWB_ADVANCED) state = NWB_ADVANCED;
    ValidityResult<float> r = eval_one<false, STATIC>(tab, q);
    if (r.hit || !r.stable) { state = NWB_TRAPPED; break; }
    double* o = nodes_out + ((size_t)e * max_steps + added) * NWB_NJ;
    for (int j = 0; j < NWB_NJ; ++j) {
      o[j] = q[j];
      cur[j] = q[j];
    }
    ++added;
  }
  n_added[e] = added;
  status[e] = (state == NWB_ADVANCED) ? -1 : state;     // -1: step budget exhausted before REACHED / TRAPPED
}

// ---- connect inside the pooled forest (RRT driver): one thread per query --------------------------------
// Same walk as connect_kernel, but the nodes go straight into tree_b of the query's forest slot and the
// end-of-iteration bookkeeping (success / connector / bridge node, tree swap, active count) happens on
// the device, so the host never has to synchronise inside an RRT iteration.
template <bool STATIC>
__global__ void __launch_bounds__(32)
connect_forest_kernel(const __grid_constant__ PlanConst pc, const __grid_constant__ ValidityTables tab, ForestView F, RrtState S,
                      int nq, const double* __restrict__ ma_pts, int n_pts) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nq || !S.active[k]) return;
  if (S.walking[k]) {
    const int t = S.tree_b[k];
    const bool start_tree = (t & 1) == 0;
    const MaView ma{ma_pts ? ma_pts + (size_t)k * n_pts * 6 : nullptr, n_pts};
    float* soa = F.soa + (size_t)t * NWB_NJ * F.cap;
    double* aos = F.aos + (size_t)t * F.cap * NWB_NJ;
    int32_t* pred = F.pred + (size_t)t * F.cap;
    int32_t* hand = ma.pts ? F.hand + (size_t)t * F.cap : nullptr;
    int size = F.size[t];
    double cur[NWB_NJ], tgt[NWB_NJ], q[NWB_NJ];
    for (int j = 0; j < NWB_NJ; ++j) {
      cur[j] = S.cstart[(size_t)k * NWB_NJ + j];
      tgt[j] = S.q_connect[(size_t)k * NWB_NJ + j];
    }
    int cur_idx = S.idx_b[k];
    int state = NWB_ADVANCED;
    int cur_hand = 0, connect_hand = 0;
    if (ma.pts) {
      cur_hand = hand[cur_idx];
      connect_hand = S.connect_hand[k];
      // the trees may only meet with the hand indices in order (RP:917-922)
      if ((start_tree && cur_hand > connect_hand) || (!start_tree && cur_hand < connect_hand)) state = NWB_TRAPPED;
    }
    while (state == NWB_ADVANCED) {
      state = generate_q_new(cur, tgt, pc, q);
      int ds = enforce_ds_constraints(q, pc, nullptr);
      if (ds == NWB_TRAPPED) { state = NWB_TRAPPED; break; }
      if (ds == NWB_ADVANCED) state = NWB_ADVANCED;
      if (ma.pts) {                              // enforce_MA_Constraints (RP:951-972)
        const int mas = enforce_ma_constraints(ma, start_tree, cur_hand, cur, q, pc);
        if (mas == NWB_TRAPPED) { state = NWB_TRAPPED; break; }
        if (mas == NWB_ADVANCED) state = NWB_ADVANCED;
        if (cur_hand == connect_hand && state != NWB_REACHED) { state = NWB_TRAPPED; break; }   // RP:968
      }
      ValidityResult<float> r = eval_one<false, STATIC>(tab, q);
      if (r.hit || !r.stable) { state = NWB_TRAPPED; break; }
      if (size >= F.cap) {                       // tree full: give the query up and report it
        atomicExch(S.overflow, 1);
        state = -1;
        break;
      }
      for (int j = 0; j < NWB_NJ; ++j) {         // addConfigtoTree (RP:749-783)
        aos[(size_t)size * NWB_NJ + j] = q[j];
        soa[(size_t)j * F.cap + size] = (float)q[j];
      }
      pred[size] = cur_idx;
      const int new_hand = ma.pts ? next_hand_index(ma, start_tree, cur_hand) : 0;
      if (ma.pts) hand[size] = new_hand;
      if (state == NWB_REACHED) {
        S.bridge[k] = cur_idx;                   // q_near = current_q_near (RP:1000-1001)
      } else {
        cur_idx = size;                          // current_q_near = tree.nodes.back() (RP:1004)
        cur_hand = new_hand;
        for (int j = 0; j < NWB_NJ; ++j) cur[j] = q[j];
      }
      ++size;
    }
    F.size[t] = size;
    if (state == NWB_REACHED || state == -1) {
      S.success[k] = state == NWB_REACHED;
      S.connector[k] = S.swap[k] ? 2 : 1;
      S.active[k] = 0;
      atomicSub(S.n_active, 1);
      return;
    }
  }
  S.swap[k] ^= 1;                                // trees swap every iteration regardless of outcome (RP:1125-1192)
}

// ---- the whole RRT-Connect query in ONE launch: one warp per query -------------------------------------------------------
// The lock-step driver above pays 7-8 launches per iteration and makes every query wait for the slowest
// one.  Here a warp owns a query from setStartGoalConfigs' trees to the connecting node: the nearest-
// neighbour scans use all 32 lanes, the serial steps (steer, projections, validity, the connect walk)
// run on lane 0 with exactly the device functions of the kernels above, so trees, iteration counts and
// bridge nodes are the same as the lock-step path's (and the oracle's).  Queries finish independently.
__device__ __forceinline__ bool lex_less_di(double da, int ia, double db, int ib) {
  return (da < db) || (da == db && (unsigned)ia < (unsigned)ib);
}

// findNearestNeighbour (RP:336-419) over tree t (n nodes) for the query in shared memory; every lane returns the winner
__device__ __forceinline__ int warp_nearest(const ForestView& F, int t, int n, const double* qd, const float* qf, int lane) {
  const float* soa = F.soa + (size_t)t * NWB_NJ * F.cap;
  const double* aos = F.aos + (size_t)t * F.cap * NWB_NJ;
  double best = 10000.0;
  int besti = -1;
  float thr = 3.0e38f;
  for (int base = 0; base < n; base += 32) {
    // The filter threshold is shared by the warp: whatever exact distance ANY lane has found bounds the answer, so every
    // lane filters with the tightest one (positive floats order like their bit patterns: one REDUX per 32 nodes).  With
    // lane-local thresholds some lane met a new local minimum in ~25 of the 31 steps of a 1000-node scan, and each such
    // step stalls the warp behind one lane's exact fp64 distance; with the shared threshold it is ~5 steps.
    thr = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(thr)));
    const int node = base + lane;
    float sf = 3.0e38f;
    if (node < n) {
      sf = 0.f;
#pragma unroll
      for (int j = 0; j < NWB_NJ; ++j) {
        float d = qf[j] - soa[(size_t)j * F.cap + node];
        sf = fmaf(d, d, sf);
      }
    }
    if (node < n && sf <= thr) {
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
    if (lex_less_di(d2, i2, best, besti)) { best = d2; besti = i2; }
  }
  return besti;
}

// ---- one configuration evaluated by a whole warp ---------------------------------------------------------------------
// The RRT kernels evaluate ONE configuration at a time (a step of a walk).  Every lane runs the forward kinematics,
// the centre of mass and the support test redundantly (same instruction stream, no extra time), then the 108
// collision pairs and the (link, box) tests are dealt out over the 32 lanes and the verdict is a warp vote.  Same
// arithmetic per pair as evaluate_config (seg_seg_dist2 & co. on the same point table), so the verdict is the same.
struct PairFlat {
  int off_a, off_b;
  float a, inv_a, e, inv_e, rsum2;
  int type;                       // 0 capsule-capsule, 1 capsule-sphere, 2 sphere-sphere
};
#define NWB_F_CC(SA, LA, ILA, SB, LB, ILB, RS, RS2) {SA, SB, LA, ILA, LB, ILB, RS2, 0},
#define NWB_F_CS(SA, LA, ILA, SB, LB, ILB, RS, RS2) {SA, SB, LA, ILA, LB, ILB, RS2, 1},
#define NWB_F_SS(SA, LA, ILA, SB, LB, ILB, RS, RS2) {SA, SB, LA, ILA, LB, ILB, RS2, 2},
__device__ const PairFlat kFlatPairsGroup[] = {NWB_PAIRS_CC_GROUP(NWB_F_CC) NWB_PAIRS_CS_GROUP(NWB_F_CS) NWB_PAIRS_SS_GROUP(NWB_F_SS)};
#undef NWB_F_CC
#undef NWB_F_CS
#undef NWB_F_SS
constexpr int kNumFlatPairsGroup = (int)(sizeof(kFlatPairsGroup) / sizeof(PairFlat));

// q: the configuration, identical on every lane (shared memory).  Returns the verdict on every lane.
__device__ __forceinline__ bool config_valid_warp(const ValidityTables& tab, const double* q, int lane) {
  float qf[NWB_NJ];
#pragma unroll
  for (int j = 0; j < NWB_NJ; ++j) qf[j] = (float)q[j];
  LocalStore st;
  ValidityVisitor<float, LocalStore, true> V(st);
  { NWB_MODEL_FK_WALK(float, XF<float>, qf, V) }
  // static stability (as evaluate_config, wedge form)
  const XF<float>& L = V.lsole;
  Wedge<float> w;
  w.init(tab.foot_right[0][0] - V.comx, tab.foot_right[0][1] - V.comy);
#pragma unroll
  for (int k = 1; k < nwbm::kFootN; ++k) w.add(tab.foot_right[k][0] - V.comx, tab.foot_right[k][1] - V.comy);
  const float ox = L.p[0] - V.comx, oy = L.p[1] - V.comy;
#pragma unroll
  for (int k = 0; k < nwbm::kFootN; ++k) {
    const float x = tab.foot_left[k][0], y = tab.foot_left[k][1];
    w.add(fmaf(L.r[1], y, fmaf(L.r[0], x, ox)), fmaf(L.r[4], y, fmaf(L.r[3], x, oy)));
  }
  const bool stable = w.surrounded;
  bool hit = false;
  // self collision: pairs lane, lane + 32, ...
  for (int p = lane; p < kNumFlatPairsGroup; p += 32) {
    const PairFlat pf = kFlatPairsGroup[p];
    float pa[3], pb[3];
    load3(st, pf.off_a, pa);
    load3(st, pf.off_b, pb);
    float d2;
    if (pf.type == 2) {
      d2 = point_point_dist2(pa, pb);
    } else {
      float qa[3], d1[3];
      load3(st, pf.off_a + 3, qa);
      d1[0] = qa[0] - pa[0]; d1[1] = qa[1] - pa[1]; d1[2] = qa[2] - pa[2];
      if (pf.type == 1) {
        d2 = seg_point_dist2(pa, d1, pf.inv_a, pb);
      } else {
        float qb[3], dd[3];
        load3(st, pf.off_b + 3, qb);
        dd[0] = qb[0] - pb[0]; dd[1] = qb[1] - pb[1]; dd[2] = qb[2] - pb[2];
        d2 = seg_seg_dist2(pa, d1, pf.a, pf.inv_a, pb, dd, pf.e, pf.inv_e);
      }
    }
    hit = hit || (d2 < pf.rsum2);
  }
  // environment: (link, box) tests lane, lane + 32, ...
  const int n_env = tab.n_env_links * tab.n_boxes;
  for (int e = lane; e < n_env; e += 32) {
    const int b = e / tab.n_env_links, l = e - b * tab.n_env_links;
    const SceneBox& bx = tab.boxes[b];
    const EnvLink& el = tab.env_links[l];
    const float H[3] = {bx.h[0], bx.h[1], bx.h[2]};
    float pa[3], A[3];
    load3(st, el.off, pa);
    const float rx = pa[0] - bx.c[0], ry = pa[1] - bx.c[1], rz = pa[2] - bx.c[2];
    A[0] = fmaf(bx.r[6], rz, fmaf(bx.r[3], ry, bx.r[0] * rx));
    A[1] = fmaf(bx.r[7], rz, fmaf(bx.r[4], ry, bx.r[1] * rx));
    A[2] = fmaf(bx.r[8], rz, fmaf(bx.r[5], ry, bx.r[2] * rx));
    float d2;
    if (el.is_sphere) {
      const float ex = box_excess(A[0], H[0]), ey = box_excess(A[1], H[1]), ez = box_excess(A[2], H[2]);
      d2 = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
    } else {
      float qa[3], D[3];
      load3(st, el.off + 3, qa);
      const float dx = qa[0] - pa[0], dy = qa[1] - pa[1], dz = qa[2] - pa[2];
      D[0] = fmaf(bx.r[6], dz, fmaf(bx.r[3], dy, bx.r[0] * dx));
      D[1] = fmaf(bx.r[7], dz, fmaf(bx.r[4], dy, bx.r[1] * dx));
      D[2] = fmaf(bx.r[8], dz, fmaf(bx.r[5], dy, bx.r[2] * dx));
      d2 = seg_box_dist2(A, D, H);
    }
    hit = hit || (d2 < el.radius * el.radius);
  }
  // large scene: the whole warp walks the hierarchy with the (identical) bounding box of the robot; at a leaf the
  // links are dealt out over the lanes
  if (tab.scene.n_nodes > 0) {
    float rlo[3] = {1e30f, 1e30f, 1e30f}, rhi[3] = {-1e30f, -1e30f, -1e30f};
    for (int l = 0; l < tab.n_env_links; ++l) {
      const EnvLink& el = tab.env_links[l];
      const float rr = el.radius + 1.0e-4f;
      float p[3];
      load3(st, el.off, p);
#pragma unroll
      for (int k = 0; k < 3; ++k) { rlo[k] = fminf(rlo[k], p[k] - rr); rhi[k] = fmaxf(rhi[k], p[k] + rr); }
      if (!el.is_sphere) {
        load3(st, el.off + 3, p);
#pragma unroll
        for (int k = 0; k < 3; ++k) { rlo[k] = fminf(rlo[k], p[k] - rr); rhi[k] = fmaxf(rhi[k], p[k] + rr); }
      }
    }
    int node = 0;
    while (node < tab.scene.n_nodes) {
      const SceneNode nd = tab.scene.nodes[node];
      const bool apart = nd.lo[0] > rhi[0] || rlo[0] > nd.hi[0] || nd.lo[1] > rhi[1] || rlo[1] > nd.hi[1] || nd.lo[2] > rhi[2] ||
                         rlo[2] > nd.hi[2];
      if (apart) { node = nd.skip; continue; }
      ++node;
      if (nd.box < 0) continue;
      const SceneBox bx = tab.scene.boxes[nd.box];
      for (int l = lane; l < tab.n_env_links; l += 32) {
        const EnvLink& el = tab.env_links[l];
        const float H[3] = {bx.h[0], bx.h[1], bx.h[2]};
        float pa[3], A[3];
        load3(st, el.off, pa);
        const float rx = pa[0] - bx.c[0], ry = pa[1] - bx.c[1], rz = pa[2] - bx.c[2];
        A[0] = fmaf(bx.r[6], rz, fmaf(bx.r[3], ry, bx.r[0] * rx));
        A[1] = fmaf(bx.r[7], rz, fmaf(bx.r[4], ry, bx.r[1] * rx));
        A[2] = fmaf(bx.r[8], rz, fmaf(bx.r[5], ry, bx.r[2] * rx));
        float d2;
        if (el.is_sphere) {
          const float ex = box_excess(A[0], H[0]), ey = box_excess(A[1], H[1]), ez = box_excess(A[2], H[2]);
          d2 = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
        } else {
          float qa[3], D[3];
          load3(st, el.off + 3, qa);
          const float dx = qa[0] - pa[0], dy = qa[1] - pa[1], dz = qa[2] - pa[2];
          D[0] = fmaf(bx.r[6], dz, fmaf(bx.r[3], dy, bx.r[0] * dx));
          D[1] = fmaf(bx.r[7], dz, fmaf(bx.r[4], dy, bx.r[1] * dx));
          D[2] = fmaf(bx.r[8], dz, fmaf(bx.r[5], dy, bx.r[2] * dx));
          d2 = seg_box_dist2(A, D, H);
        }
        hit = hit || (d2 < el.radius * el.radius);
      }
    }
  }
  hit = __any_sync(0xffffffffu, hit);
  return !hit && stable;
}

// one out-of-line copy of the full validity evaluation for the RRT kernels (they call it from four places)
// q lives in shared memory and is the same for every lane; all 32 lanes call this together.  With the compiled-in
// pair set the warp shares the work; the table-driven set (srdf_trim = 0) is evaluated by lane 0 as before.
template <bool STATIC>
static __device__ __noinline__ bool config_valid(const ValidityTables& tab, const double* q, int lane) {
  if constexpr (STATIC) {
    return config_valid_warp(tab, q, lane);
  } else {
    int ok = 0;
    if (lane == 0) {
      ValidityResult<float> v = eval_one<false, false>(tab, q);
      ok = !v.hit && v.stable;
    }
    return __shfl_sync(0xffffffffu, ok, 0) != 0;
  }
}

struct RrtOutcome {
  int iterations, success, connector, bridge, overflow;
};

// solveQuery's loop (RP:1067-1198) for query k, executed by one warp.  size[0/1] = node counts of the start / goal tree
// (kept in registers of every lane, updated as nodes are appended); qd / qf = 21-element shared scratch of the warp
// (query of the nearest-neighbour scans), qv = the configuration handed to the warp-wide validity evaluation.
// The serial arithmetic (steer, projections, tree bookkeeping) runs on lane 0; the scans and the validity
// evaluation use all lanes, so every lane walks the same control flow (decisions are broadcast from lane 0).
template <bool STATIC>
__device__ __forceinline__ RrtOutcome rrt_query_loop(const PlanConst& pc, const ValidityTables& tab, const ForestView& F, int k, int lane,
                                                     uint64_t seed, const double* __restrict__ db, int n_db, int max_iter,
                                                     const MaView& ma, double* qd, float* qf, double* qv, WarpSteer& ws,
                                                     int (&size)[2], int32_t* overflow_flag) {
  const unsigned FULL = 0xffffffffu;
  int swap = 0, done = 0, success = 0, connector = 0, bridge = 0, it = 0;
  for (; it < max_iter && !done; ++it) {
    // getRandomStableConfig (RP:316-332) with the injected stream
    uint32_t r[4];
    philox4x32_10((uint32_t)it, 0u, 0u, kStreamPlanner, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u = u01(r[0]);
    const int row = (pc.quirks & NWB_QUIRK_DB_INDEX_FLOOR) ? (int)floor(0.0 + (double)(n_db - 1 - 0) * u) : (int)floor((double)n_db * u);
    const int ta = 2 * k + swap, tb = 2 * k + 1 - swap;
    if (lane < NWB_NJ) {
      const double v = db[(size_t)row * NWB_NJ + lane];
      qd[lane] = v;
      qf[lane] = (float)v;
    }
    __syncwarp();
    // ---- extendTree (RP:787-884)
    const int idx_a = warp_nearest(F, ta, size[ta & 1], qd, qf, lane);
    const bool a_start = (ta & 1) == 0;
    int near_hand = 0, go = 0;
    if (lane < NWB_NJ) ws.near[lane] = F.aos[((size_t)ta * F.cap + idx_a) * NWB_NJ + lane];
    __syncwarp();
    int st = warp_steer_project(pc, ws, qd, qv, lane);    // generate_q_new + enforce_DS_Constraints: qv = the candidate node
    if (st != NWB_TRAPPED && ma.pts) {                    // enforce_MA_Constraints (RP:829-837), lane 0
      if (lane == 0) {
        near_hand = F.hand[(size_t)ta * F.cap + idx_a];
        const int mas = enforce_ma_constraints(ma, a_start, near_hand, ws.near, qv, pc);
        if (mas == NWB_TRAPPED) st = NWB_TRAPPED;
        else if (mas == NWB_ADVANCED) st = NWB_ADVANCED;
      }
      st = __shfl_sync(FULL, st, 0);
      __syncwarp();
    }
    go = st != NWB_TRAPPED;
    int walking = 0, connect_hand = 0;
    if (go) {
      const bool valid = config_valid<STATIC>(tab, qv, lane);         // isConfigValid (RP:846), all lanes
      if (valid) {
        const int base = size[ta & 1];
        if (base >= F.cap) {
          if (lane == 0) atomicExch(overflow_flag, 1);
          done = 2;
        } else {
          if (lane == 0) {                                // addConfigtoTree (RP:749-783)
            for (int j = 0; j < NWB_NJ; ++j) {
              const double v = qv[j];
              F.aos[((size_t)ta * F.cap + base) * NWB_NJ + j] = v;
              F.soa[(size_t)ta * NWB_NJ * F.cap + (size_t)j * F.cap + base] = (float)v;
              qd[j] = v;                                  // q_connect of this iteration's connectTree
              qf[j] = (float)v;
            }
            F.pred[(size_t)ta * F.cap + base] = idx_a;
            if (ma.pts) {
              connect_hand = next_hand_index(ma, a_start, near_hand);
              F.hand[(size_t)ta * F.cap + base] = connect_hand;
            }
          }
          walking = 1;
          size[ta & 1] += 1;
        }
      }
    }
    __syncwarp();
    if (walking) {
      // ---- connectTree (RP:890-1034)
      const int idx_b = warp_nearest(F, tb, size[tb & 1], qd, qf, lane);
      const bool b_start = (tb & 1) == 0;
      float* soa = F.soa + (size_t)tb * NWB_NJ * F.cap;
      double* aos = F.aos + (size_t)tb * F.cap * NWB_NJ;
      int32_t* pred = F.pred + (size_t)tb * F.cap;
      int32_t* hand = ma.pts ? F.hand + (size_t)tb * F.cap : nullptr;
      int sz = size[tb & 1];
      // the walk's current node lives in ws.near (shared); its index / hand index on lane 0
      int cur_idx = idx_b, state = NWB_ADVANCED, cur_hand = 0;
      if (lane < NWB_NJ) ws.near[lane] = aos[(size_t)idx_b * NWB_NJ + lane];
      if (lane == 0 && ma.pts) {
        cur_hand = hand[cur_idx];
        if ((b_start && cur_hand > connect_hand) || (!b_start && cur_hand < connect_hand)) state = NWB_TRAPPED;   // RP:917-922
      }
      state = __shfl_sync(FULL, state, 0);
      __syncwarp();
      while (state == NWB_ADVANCED) {
        state = warp_steer_project(pc, ws, qd, qv, lane);  // qd = q_connect; qv = the candidate node
        if (state != NWB_TRAPPED && ma.pts) {             // RP:951-972, lane 0
          if (lane == 0) {
            bool proceed = true;
            const int mas = enforce_ma_constraints(ma, b_start, cur_hand, ws.near, qv, pc);
            if (mas == NWB_TRAPPED) proceed = false;
            else {
              if (mas == NWB_ADVANCED) state = NWB_ADVANCED;
              if (cur_hand == connect_hand && state != NWB_REACHED) proceed = false;
            }
            if (!proceed) state = NWB_TRAPPED;
          }
          state = __shfl_sync(FULL, state, 0);            // ADVANCED / REACHED with a candidate in qv, TRAPPED otherwise
          __syncwarp();
        }
        if (state == NWB_TRAPPED) break;
        const bool valid = config_valid<STATIC>(tab, qv, lane);
        if (!valid) { state = NWB_TRAPPED; break; }
        if (sz >= F.cap) {
          if (lane == 0) atomicExch(overflow_flag, 1);
          state = -1;
          break;
        }
        if (lane == 0) {
          for (int j = 0; j < NWB_NJ; ++j) {
            const double v = qv[j];
            aos[(size_t)sz * NWB_NJ + j] = v;
            soa[(size_t)j * F.cap + sz] = (float)v;
          }
          pred[sz] = cur_idx;
          const int new_hand = ma.pts ? next_hand_index(ma, b_start, cur_hand) : 0;
          if (ma.pts) hand[sz] = new_hand;
          if (state == NWB_REACHED) {
            bridge = cur_idx;                             // RP:1000-1001
          } else {
            cur_idx = sz;                                 // RP:1004
            cur_hand = new_hand;
          }
        }
        if (state != NWB_REACHED && lane < NWB_NJ) ws.near[lane] = qv[lane];   // RP:1004: the next step starts from the projected node
        ++sz;
        __syncwarp();
      }
      size[tb & 1] = sz;
      if (state == NWB_REACHED || state == -1) {
        success = state == NWB_REACHED;
        connector = swap ? 2 : 1;
        done = 1;
      }
    }
    swap ^= 1;                                            // RP:1125-1192: every iteration, whatever the outcome
  }
  RrtOutcome o;                                           // bridge lives on lane 0: give it to every lane
  o.iterations = it;
  o.success = success;
  o.connector = connector;
  o.bridge = __shfl_sync(FULL, bridge, 0);
  o.overflow = done == 2 || (done == 1 && !success);
  return o;
}

template <bool STATIC>
__global__ void __launch_bounds__(32)
rrt_persistent_kernel(const __grid_constant__ PlanConst pc, const __grid_constant__ ValidityTables tab, ForestView F, RrtState S,
                      const double* __restrict__ db, int n_db, const uint64_t* __restrict__ seeds, int max_iter, int nq,
                      const double* __restrict__ ma_pts, int n_pts) {
  const int k = blockIdx.x;
  const int lane = threadIdx.x;
  if (k >= nq || !S.active[k]) return;
  __shared__ double qd[NWB_NJ];          // current query of the nearest-neighbour scan (q_rand, then q_connect)
  __shared__ float qf[NWB_NJ];
  __shared__ double qv[NWB_NJ];          // configuration under the warp-wide validity evaluation
  __shared__ WarpSteer ws;               // scratch of the warp-cooperative steer + projection
  const MaView ma{ma_pts ? ma_pts + (size_t)k * n_pts * 6 : nullptr, n_pts};
  int size[2] = {F.size[2 * k], F.size[2 * k + 1]};      // kept in registers of every lane
  const RrtOutcome o = rrt_query_loop<STATIC>(pc, tab, F, k, lane, seeds[k], db, n_db, max_iter, ma, qd, qf, qv, ws, size, S.overflow);
  if (lane == 0) {
    F.size[2 * k] = size[0];
    F.size[2 * k + 1] = size[1];
    S.iterations[k] = o.iterations;
    S.success[k] = o.success;
    S.connector[k] = o.connector;
    S.bridge[k] = o.bridge;
    S.active[k] = 0;
  }
}

// Fused form: the query's roots are validated and inserted here (setStartGoalConfigs, RP:227-302), the loop runs,
// and on success the raw path (writePath's extraction, RP:1303-1392) is written into a shared row pool - so the
// host issues ONE upload, ONE launch and ONE download per call.  res[k] = {start_valid, goal_valid, success,
// iterations, connector, bridge, nodes_start, nodes_goal, path_len, path_offset (-1: pool exhausted), overflow, 0}.
struct RrtIo {
  const double* starts;     // [nq][21]
  const double* goals;      // [nq][21]
  const uint64_t* seeds;    // [nq]
  int32_t* res;             // [nq][12]
  double* path_rows;        // [path_cap][21]
  unsigned long long* path_used;
  long long path_cap;
};

template <bool STATIC>
__global__ void __launch_bounds__(32)
rrt_fused_kernel(const __grid_constant__ PlanConst pc, const __grid_constant__ ValidityTables tab, ForestView F, RrtIo io,
                 const double* __restrict__ db, int n_db, int max_iter, int nq, const double* __restrict__ ma_pts, int n_pts,
                 int32_t* __restrict__ overflow_flag) {
  const int k = blockIdx.x;
  const int lane = threadIdx.x;
  if (k >= nq) return;
  __shared__ double qd[NWB_NJ];
  __shared__ float qf[NWB_NJ];
  __shared__ double qv[NWB_NJ];
  __shared__ WarpSteer ws;
  const MaView ma{ma_pts ? ma_pts + (size_t)k * n_pts * 6 : nullptr, n_pts};
  int32_t* res = io.res + (size_t)k * 12;
  const int ts = 2 * k, tg = 2 * k + 1;
  // ---- setStartGoalConfigs: both roots must be valid (RP:241-265)
  if (lane < NWB_NJ) qv[lane] = io.starts[(size_t)k * NWB_NJ + lane];
  __syncwarp();
  const int sv = config_valid<STATIC>(tab, qv, lane);
  __syncwarp();
  if (lane < NWB_NJ) qv[lane] = io.goals[(size_t)k * NWB_NJ + lane];
  __syncwarp();
  const int gv = config_valid<STATIC>(tab, qv, lane);
  const int ok = sv && gv;
  if (lane == 0) {
    for (int j = 0; j < 12; ++j) res[j] = 0;
    res[0] = sv;
    res[1] = gv;
    res[9] = -1;
    if (ok) {                                             // roots: index 0, predecessor 0 (RP:236-258)
      for (int side = 0; side < 2; ++side) {
        const int t = side ? tg : ts;
        const double* src = (side ? io.goals : io.starts) + (size_t)k * NWB_NJ;
        for (int j = 0; j < NWB_NJ; ++j) {
          F.aos[(size_t)t * F.cap * NWB_NJ + j] = src[j];
          F.soa[(size_t)t * NWB_NJ * F.cap + (size_t)j * F.cap] = (float)src[j];
        }
        F.pred[(size_t)t * F.cap] = 0;
        if (ma.pts) F.hand[(size_t)t * F.cap] = side ? n_pts - 1 : 0;   // RP:274-282
      }
    }
  }
  __syncwarp();
  if (!ok) return;
  int size[2] = {1, 1};
  const RrtOutcome o = rrt_query_loop<STATIC>(pc, tab, F, k, lane, io.seeds[k], db, n_db, max_iter, ma, qd, qf, qv, ws, size, overflow_flag);
  // ---- results + raw path
  int ls = 0, lg = 0, idx_s = 0, idx_g = 0;
  long long off = -1;
  const int32_t* pred_s = F.pred + (size_t)ts * F.cap;
  const int32_t* pred_g = F.pred + (size_t)tg * F.cap;
  if (lane == 0) {
    F.size[ts] = size[0];
    F.size[tg] = size[1];
    res[2] = o.success;
    res[3] = o.iterations;
    res[4] = o.success ? o.connector : 0;
    res[5] = o.success ? o.bridge : 0;
    res[6] = size[0];
    res[7] = size[1];
    res[10] = o.overflow;
    if (o.success) {
      if (o.connector == 1) { idx_s = size[0] - 1; idx_g = o.bridge; }       // RP:1317-1327
      else { idx_s = o.bridge; idx_g = size[1] - 1; }
      ls = lg = 1;
      for (int i = idx_s; i > 0; i = pred_s[i]) ++ls;
      for (int i = idx_g; i > 0; i = pred_g[i]) ++lg;
      const unsigned long long o0 = atomicAdd(io.path_used, (unsigned long long)(ls + lg));
      if ((long long)(o0 + ls + lg) <= io.path_cap) off = (long long)o0;
      res[8] = ls + lg;
      res[9] = (int32_t)off;
    }
  }
  __syncwarp();
  if (!o.success) return;
  ls = __shfl_sync(0xffffffffu, ls, 0);
  lg = __shfl_sync(0xffffffffu, lg, 0);
  idx_s = __shfl_sync(0xffffffffu, idx_s, 0);
  idx_g = __shfl_sync(0xffffffffu, idx_g, 0);
  off = __shfl_sync(0xffffffffu, off, 0);
  if (off < 0) return;                                    // pool exhausted: the host extracts this path separately
  double* out = io.path_rows + off * NWB_NJ;
  int i = idx_s;
  for (int pos = ls - 1; pos >= 0; --pos) {               // start branch, reversed into place
    if (lane < NWB_NJ) out[(size_t)pos * NWB_NJ + lane] = F.aos[((size_t)ts * F.cap + i) * NWB_NJ + lane];
    i = pred_s[i];
  }
  i = idx_g;
  for (int pos = 0; pos < lg; ++pos) {                    // goal branch as walked
    if (lane < NWB_NJ) out[(size_t)(ls + pos) * NWB_NJ + lane] = F.aos[((size_t)tg * F.cap + i) * NWB_NJ + lane];
    i = pred_g[i];
  }
}

// ---- edge check: (edge, way-point) -> collision-only validity ------------------------------------------
// ONE launch.  A block owns whole edges (as many as fill two passes of its 256 threads: 5 edges x 102 way-points), so the
// first colliding way-point of an edge is a shared-memory minimum and the block writes ok / first_bad itself - no
// memset before, no finishing kernel after, no global atomics.  The end points of the block's edges are staged in shared
// memory once (start + step width per joint, fp64 as interpolateWaypoints computes them, RP:1758-1771); each thread
// forms its way-point and evaluates the whole-robot collision predicate in registers.
constexpr int EC_THREADS = 256;     // two resident blocks per SM: one block's epilogue overlaps the other's evaluation
constexpr int EC_SLOTS = 512;       // (edge, way-point) slots a block aims at: two passes of its threads
constexpr int EC_MAX_EDGES = 64;    // edges per block (reached only for k + 2 <= 8 way-points per edge)

// EARLY_EXIT: a way-point whose edge already has an earlier colliding way-point on record is skipped;
// the warp votes (ballot) and moves on together, so no lane idles through a full evaluation for an
// edge that is already lost.  Results (ok, first_bad) are identical to the full evaluation.
template <bool STATIC, bool EARLY_EXIT>
__global__ void __launch_bounds__(EC_THREADS, 2)
edge_check_kernel(const __grid_constant__ ValidityTables tab, const double* __restrict__ qa, const double* __restrict__ qb,
                  int64_t n_edges, int k, int edges_per_block, uint8_t* __restrict__ ok, int32_t* __restrict__ first_bad) {
  __shared__ double s_a[EC_MAX_EDGES][NWB_NJ], s_sw[EC_MAX_EDGES][NWB_NJ];
  __shared__ int s_bad[EC_MAX_EDGES];
  const int wp = k + 2;
  const int64_t e0 = (int64_t)blockIdx.x * edges_per_block;
  const int ne = (int)min((int64_t)edges_per_block, n_edges - e0);
  for (int t = threadIdx.x; t < ne * NWB_NJ; t += blockDim.x) {
    const int le = t / NWB_NJ, j = t - le * NWB_NJ;
    const double a = qa[(e0 + le) * NWB_NJ + j], b = qb[(e0 + le) * NWB_NJ + j];
    s_a[le][j] = a;
    s_sw[le][j] = (b - a) / (k + 1);                        // RP:1758
  }
  if (threadIdx.x < ne) s_bad[threadIdx.x] = 0x7fffffff;
  __syncthreads();
  const int total = ne * wp;
  for (int base = 0; base < total; base += blockDim.x) {  // two passes for 5 x 102 way-points; more when one edge alone is longer
    const int slot = base + threadIdx.x;
    const bool active = slot < total;
    if (__ballot_sync(0xffffffffu, active) == 0u) continue;
    const int sl = active ? slot : 0;
    const int le = sl / wp, n = sl - le * wp;
    if constexpr (EARLY_EXIT) {
      const bool needed = active && (*reinterpret_cast<volatile int*>(&s_bad[le]) > n);
      if (__ballot_sync(0xffffffffu, needed) == 0u) continue;   // the whole warp is past a known collision
    }
    double q[NWB_NJ];
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) {
      const double a = s_a[le][j];
      q[j] = (n == 0) ? a : __dadd_rn(a, __dmul_rn(s_sw[le][j], (double)n));   // RP:1763, 1771
    }
    ValidityResult<float> r = eval_one<true, STATIC>(tab, q);
    if (active && r.hit) atomicMin(&s_bad[le], n);
  }
  __syncthreads();
  if (threadIdx.x < ne) {
    const int fb = s_bad[threadIdx.x];
    const bool good = fb == 0x7fffffff;                    // no way-point collided
    if (ok) ok[e0 + threadIdx.x] = good;
    first_bad[e0 + threadIdx.x] = good ? -1 : fb;
  }
}

// ---- stable-configuration sampler ------------------------------------------------------------------------
// Sample i draws its 21 joints from Philox(seed; counter = (i, slot/4, stream)), so the accepted set
// is independent of how [first, first+count) is partitioned over launches, streams or GPUs.
#ifndef NWB_SAMPLE_MIN_BLOCKS
#define NWB_SAMPLE_MIN_BLOCKS 6      // resident 64-thread blocks per SM the register allocation aims at (168 registers;
                                     // measured 23.4 M samples/s against 21.5 at 4 blocks / 255 registers and 21.9 at 8 / 128)
#endif

// sample i -> its 21 uniform joints, LHipYawPitch = 0 (SCG:285-291)
__device__ __forceinline__ void sampler_draw(const PlanConst& pc, uint64_t seed, uint64_t i, double* q) {
#pragma unroll 1
  for (int blk = 0; blk < (NWB_NJ + 3) / 4; ++blk) {
    uint32_t r[4];
    philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)blk, kStreamSampler, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int l = 0; l < 4; ++l) {
      int j = 4 * blk + l;
      if (j < NWB_NJ) q[j] = __dadd_rn(__dmul_rn(pc.jmax[j] - pc.jmin[j], u01(r[l])), pc.jmin[j]);   // uniformReal(lo, hi)
    }
  }
  q[15] = 0.0;                                           // SCG:291
}

// Phase 1: one thread per sample.  Draw the sample, run the right-leg FK and decide whether the left-foot target
// can be reached at all (the exact test of left_leg_ik_nr).  The samples that still need the Newton-Raphson solve
// are appended to a dense work list (one atomic per warp), so that phase 2 runs warps FULL of solves instead of
// warps in which every lane waits for the one lane that needs all 100 iterations.
__global__ void __launch_bounds__(128)
sample_filter_kernel(const __grid_constant__ PlanConst pc, uint64_t seed, uint64_t first, int64_t count, int max_ik_iter,
                     uint32_t* __restrict__ work, unsigned* __restrict__ work_n, uint8_t* __restrict__ stage_out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool need = false;
  if (t < count && max_ik_iter > 0) {
    double q[NWB_NJ], rl[5];
    sampler_draw(pc, seed, first + (uint64_t)t, q);
    for (int m = 0; m < 5; ++m) rl[m] = -q[m];           // SCG:307-313
    DFrame hip, target;
    right_leg_fk(rl, hip);
    desired_left_leg_pose(hip, target);
    need = left_leg_target_feasible(target, pc.ik_eps);
  }
  if (t < count && stage_out) stage_out[t] = 0;           // phase 2 overwrites it for the samples it solves
  const unsigned mask = __ballot_sync(0xffffffffu, need);
  if (mask) {
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(work_n, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (need) work[base + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)t;
  }
}

// dense append of the lanes with `keep`: one atomic per warp; every lane of the warp must call it
__device__ __forceinline__ unsigned sampler_warp_append(bool keep, unsigned* counter) {
  const unsigned mask = __ballot_sync(0xffffffffu, keep);
  if (!mask) return 0;
  const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
}

// what follows a converged solve: joint limits, validity, compaction of the kept rows
template <bool STATIC>
__device__ __forceinline__ void sampler_finish(const PlanConst& pc, const ValidityTables& tab, double* q, const double* ll, uint64_t i,
                                               int64_t t, double* __restrict__ rows, uint64_t* __restrict__ row_index, int64_t cap,
                                               unsigned long long* __restrict__ n_out, uint8_t* __restrict__ stage_out) {
  if (!limits_ok(ll, pc.jmin + 16, pc.jmax + 16, 5)) return;              // stage stays 0 (written by the filter)
  for (int m = 0; m < 5; ++m) q[16 + m] = ll[m];
  ValidityResult<float> r = eval_one<false, STATIC>(tab, q);
  const bool feasible = !r.hit && r.stable;
  const bool keep = (pc.quirks & NWB_QUIRK_DB_KEEPS_INFEASIBLE) ? !feasible : feasible;   // SCG:452-456
  if (keep) {
    // system scope: the output arrays may live on ANOTHER GPU (peer-mapped through nwb_ipc_import): every rank of
    // a sharded generation then appends into one rank's database directly over NVLink - the gather is the store
    unsigned long long slot = atomicAdd_system(n_out, 1ull);
    if ((int64_t)slot < cap) {
      for (int j = 0; j < NWB_NJ; ++j) rows[slot * NWB_NJ + j] = q[j];
      row_index[slot] = i;
    }
  }
  if (stage_out) stage_out[t] = feasible ? 2 : 1;
}

// Phase 2 (head): one thread per work item, the first kSamplerHead Newton-Raphson iterations.  91 % of the solves that
// converge at all have done so by then; the others (and everything that will fail after all 100 iterations) hand
// their state - 5 joints and the 5x5 right singular vectors - to a second dense list, so the long tail again runs
// in full warps instead of warps where the converged lanes idle.  A lane that finds the tail list full simply keeps
// iterating here.  counters[0] = |work|, counters[1] = |tail|.
constexpr int kSamplerHead = 10;
constexpr int kTailState = 30;                           // q[5] + V[5][5], struct-of-arrays [kTailState][tail_cap]
template <bool STATIC>
__global__ void __launch_bounds__(64, NWB_SAMPLE_MIN_BLOCKS)
sample_kernel(const __grid_constant__ PlanConst pc, const __grid_constant__ ValidityTables tab, uint64_t seed, uint64_t first,
              const uint32_t* __restrict__ work, unsigned* __restrict__ counters, uint32_t* __restrict__ tail_t,
              double* __restrict__ tail_state, unsigned tail_cap, double* __restrict__ rows /*[cap][21]*/,
              uint64_t* __restrict__ row_index, int64_t cap, unsigned long long* __restrict__ n_out, uint8_t* __restrict__ stage_out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = w < (int64_t)counters[0];
  if (__ballot_sync(0xffffffffu, active) == 0) return;
  int64_t t = 0;
  uint64_t i = 0;
  double q[NWB_NJ], ll[5], V[5][5];
  DFrame target;
  int it = -1;
  const int head = pc.ik_maxiter < kSamplerHead ? pc.ik_maxiter : kSamplerHead;
  if (active) {
    t = work[w];
    i = first + (uint64_t)t;
    sampler_draw(pc, seed, i, q);
    double rl[5];
    for (int m = 0; m < 5; ++m) rl[m] = -q[m];           // SCG:307-313
    DFrame hip;
    right_leg_fk(rl, hip);
    desired_left_leg_pose(hip, target);
    for (int m = 0; m < 5; ++m) ll[m] = -rl[m];          // NCK:150 seed; every further attempt repeats the same arithmetic (quirk C5)
    nr_identity(V);
    it = left_leg_ik_nr_span(target, ll, V, 0, head, pc.ik_eps, pc.pinv_eps);
  }
  const bool more = active && it < 0 && head < pc.ik_maxiter;
  const unsigned slot = sampler_warp_append(more, counters + 1);
  if (more) {
    if (slot < tail_cap) {
      tail_t[slot] = (uint32_t)t;
#pragma unroll
      for (int m = 0; m < 5; ++m) tail_state[(size_t)m * tail_cap + slot] = ll[m];
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) tail_state[(size_t)(5 + 5 * a + b) * tail_cap + slot] = V[a][b];
      return;
    }
    it = left_leg_ik_nr_span(target, ll, V, head, pc.ik_maxiter, pc.ik_eps, pc.pinv_eps);
  }
  if (active && it >= 0) sampler_finish<STATIC>(pc, tab, q, ll, i, t, rows, row_index, cap, n_out, stage_out);
}

// Phase 2 (tail): iterations kSamplerHead .. maxiter of the solves that were still running.
template <bool STATIC>
__global__ void __launch_bounds__(64, NWB_SAMPLE_MIN_BLOCKS)
sample_tail_kernel(const __grid_constant__ PlanConst pc, const __grid_constant__ ValidityTables tab, uint64_t seed, uint64_t first,
                   const unsigned* __restrict__ counters, const uint32_t* __restrict__ tail_t, const double* __restrict__ tail_state,
                   unsigned tail_cap, double* __restrict__ rows, uint64_t* __restrict__ row_index, int64_t cap,
                   unsigned long long* __restrict__ n_out, uint8_t* __restrict__ stage_out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned queued = counters[1] < tail_cap ? counters[1] : tail_cap;
  if (w >= (int64_t)queued) return;
  const int64_t t = tail_t[w];
  const uint64_t i = first + (uint64_t)t;
  double q[NWB_NJ], ll[5], V[5][5], rl[5];
  sampler_draw(pc, seed, i, q);
  for (int m = 0; m < 5; ++m) rl[m] = -q[m];
  DFrame hip, target;
  right_leg_fk(rl, hip);
  desired_left_leg_pose(hip, target);
#pragma unroll
  for (int m = 0; m < 5; ++m) ll[m] = tail_state[(size_t)m * tail_cap + w];
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int b = 0; b < 5; ++b) V[a][b] = tail_state[(size_t)(5 + 5 * a + b) * tail_cap + w];
  const int it = left_leg_ik_nr_span(target, ll, V, kSamplerHead, pc.ik_maxiter, pc.ik_eps, pc.pinv_eps);
  if (it >= 0) sampler_finish<STATIC>(pc, tab, q, ll, i, t, rows, row_index, cap, n_out, stage_out);
}

// ---- launchers -------------------------------------------------------------------------------------------------
static unsigned steer_tail_cap(int64_t n) { return (unsigned)std::max<int64_t>(1024, n / 4); }
size_t steer_scratch_bytes(int64_t n) {      // counters | tail state [30][cap] | tail index [cap] | work [n]
  const size_t tc = steer_tail_cap(n);
  return 256 + sizeof(double) * kSteerTailState * tc + sizeof(uint32_t) * tc + sizeof(uint32_t) * (size_t)std::max<int64_t>(n, 0);
}
cudaError_t launch_steer_project(const PlanConst& pc, const double* q_near, const double* q_target, int64_t n, double* q_out,
                                 int32_t* status, int32_t* ds_status, int32_t* iters, void* scratch, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (!scratch || n < 4096 || n > (int64_t)0xffffffff) {       // small batches (the lock-step driver): one launch
    steer_project_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(pc, q_near, q_target, n, q_out, status, ds_status, iters);
    return cudaGetLastError();
  }
  const unsigned tc = steer_tail_cap(n);
  unsigned* counters = static_cast<unsigned*>(scratch);        // scratch: steer_scratch_bytes(n)
  double* tail_state = reinterpret_cast<double*>(static_cast<char*>(scratch) + 256);
  uint32_t* tail_i = reinterpret_cast<uint32_t*>(tail_state + (size_t)kSteerTailState * tc);
  uint32_t* work = tail_i + tc;
  cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned), s);
  if (e != cudaSuccess) return e;
  steer_filter_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(pc, q_near, q_target, n, q_out, status, ds_status, iters, work, counters);
  // the list sizes stay on the device: launch for the worst case, surplus blocks leave at once
  steer_solve_kernel<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(pc, q_near, q_target, work, counters, tail_i, tail_state, tc, q_out, status,
                                                             ds_status, iters);
  steer_tail_kernel<<<(tc + 63) / 64, 64, 0, s>>>(pc, q_near, q_target, counters, tail_i, tail_state, tc, q_out, status, ds_status, iters);
  return cudaGetLastError();
}

cudaError_t launch_connect(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, const double* q_start,
                           const double* q_connect, int64_t n_edges, int max_steps, double* nodes_out, int32_t* n_added,
                           int32_t* status, cudaStream_t s) {
  if (n_edges <= 0) return cudaSuccess;
  unsigned blocks = (unsigned)((n_edges + 63) / 64);
  if (static_pairs) connect_kernel<true><<<blocks, 64, 0, s>>>(pc, tab, q_start, q_connect, n_edges, max_steps, nodes_out, n_added, status);
  else connect_kernel<false><<<blocks, 64, 0, s>>>(pc, tab, q_start, q_connect, n_edges, max_steps, nodes_out, n_added, status);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(128)
pack_rows_kernel(const double* __restrict__ nodes, const int32_t* __restrict__ n_added, const int64_t* __restrict__ offset,
                 int max_steps, double* __restrict__ packed) {
  const int64_t e = blockIdx.x;
  const int k = min(max(n_added[e], 0), max_steps);
  const double* src = nodes + (size_t)e * max_steps * NWB_NJ;
  double* dst = packed + (size_t)offset[e] * NWB_NJ;
  for (int x = threadIdx.x; x < k * NWB_NJ; x += blockDim.x) dst[x] = src[x];
}
cudaError_t launch_pack_rows(const double* nodes, const int32_t* n_added, const int64_t* offset, int64_t n_edges, int max_steps,
                             double* packed, cudaStream_t s) {
  if (n_edges <= 0) return cudaSuccess;
  pack_rows_kernel<<<(unsigned)n_edges, 128, 0, s>>>(nodes, n_added, offset, max_steps, packed);
  return cudaGetLastError();
}

cudaError_t launch_connect_forest(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, ForestView F, RrtState S,
                                  int nq, const double* ma_pts, int n_pts, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  unsigned blocks = (unsigned)((nq + 31) / 32);
  if (static_pairs) connect_forest_kernel<true><<<blocks, 32, 0, s>>>(pc, tab, F, S, nq, ma_pts, n_pts);
  else connect_forest_kernel<false><<<blocks, 32, 0, s>>>(pc, tab, F, S, nq, ma_pts, n_pts);
  return cudaGetLastError();
}

cudaError_t launch_rrt_persistent(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, ForestView F, RrtState S,
                                  const double* db, int n_db, const uint64_t* seeds, int max_iter, int nq, const double* ma_pts,
                                  int n_pts, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  if (static_pairs) rrt_persistent_kernel<true><<<nq, 32, 0, s>>>(pc, tab, F, S, db, n_db, seeds, max_iter, nq, ma_pts, n_pts);
  else rrt_persistent_kernel<false><<<nq, 32, 0, s>>>(pc, tab, F, S, db, n_db, seeds, max_iter, nq, ma_pts, n_pts);
  return cudaGetLastError();
}

cudaError_t launch_rrt_fused(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, ForestView F, const double* starts,
                             const double* goals, const uint64_t* seeds, int32_t* res, double* path_rows,
                             unsigned long long* path_used, long long path_cap, const double* db, int n_db, int max_iter, int nq,
                             const double* ma_pts, int n_pts, int32_t* overflow_flag, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  RrtIo io{starts, goals, seeds, res, path_rows, path_used, path_cap};
  if (static_pairs) rrt_fused_kernel<true><<<nq, 32, 0, s>>>(pc, tab, F, io, db, n_db, max_iter, nq, ma_pts, n_pts, overflow_flag);
  else rrt_fused_kernel<false><<<nq, 32, 0, s>>>(pc, tab, F, io, db, n_db, max_iter, nq, ma_pts, n_pts, overflow_flag);
  return cudaGetLastError();
}

cudaError_t launch_edge_check(const ValidityTables& tab, bool static_pairs, const double* qa, const double* qb, int64_t n_edges,
                              int k, uint8_t* ok, int32_t* first_bad, cudaStream_t s) {
  if (n_edges <= 0) return cudaSuccess;
  const int wp = k + 2;
  const int epb = std::max(1, std::min(EC_MAX_EDGES, EC_SLOTS / wp));        // whole edges per block
  const int threads = std::max(32, std::min(EC_THREADS, (epb * wp + 31) / 32 * 32));
  const unsigned blocks = (unsigned)((n_edges + epb - 1) / epb);
  static const bool full = getenv("NWB_EDGE_FULL") != nullptr;    // roofline runs: evaluate every way-point
  if (static_pairs) {
    if (full) edge_check_kernel<true, false><<<blocks, threads, 0, s>>>(tab, qa, qb, n_edges, k, epb, ok, first_bad);
    else edge_check_kernel<true, true><<<blocks, threads, 0, s>>>(tab, qa, qb, n_edges, k, epb, ok, first_bad);
  } else {
    if (full) edge_check_kernel<false, false><<<blocks, threads, 0, s>>>(tab, qa, qb, n_edges, k, epb, ok, first_bad);
    else edge_check_kernel<false, true><<<blocks, threads, 0, s>>>(tab, qa, qb, n_edges, k, epb, ok, first_bad);
  }
  return cudaGetLastError();
}

// the sampler works through its range in chunks of at most kSamplerChunk samples, so the scratch stays bounded
constexpr int64_t kSamplerChunk = (int64_t)1 << 22;
static unsigned sampler_tail_cap(int64_t chunk) { return (unsigned)std::max<int64_t>(1024, chunk / 4); }
size_t sampler_scratch_bytes(int64_t count) {
  const int64_t chunk = std::min(count, kSamplerChunk);
  const size_t tc = sampler_tail_cap(chunk);
  return 256 + sizeof(double) * kTailState * tc + sizeof(uint32_t) * tc + sizeof(uint32_t) * (size_t)chunk;
}
cudaError_t launch_sample(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, uint64_t seed, uint64_t first,
                          int64_t count, int max_ik_iter, double* rows, uint64_t* row_index, int64_t cap,
                          unsigned long long* n_out, uint8_t* stage_out, void* scratch, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const int64_t chunk_max = std::min(count, kSamplerChunk);
  const unsigned tc = sampler_tail_cap(chunk_max);
  unsigned* counters = static_cast<unsigned*>(scratch);
  double* tail_state = reinterpret_cast<double*>(static_cast<char*>(scratch) + 256);
  uint32_t* tail_t = reinterpret_cast<uint32_t*>(tail_state + (size_t)kTailState * tc);
  uint32_t* work = tail_t + tc;
  for (int64_t off = 0; off < count; off += kSamplerChunk) {
    const int64_t n = std::min(kSamplerChunk, count - off);
    uint8_t* st = stage_out ? stage_out + off : nullptr;
    cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
    sample_filter_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(pc, seed, first + (uint64_t)off, n, max_ik_iter, work, counters, st);
    // the list sizes stay on the device: launch for the worst case, surplus blocks leave at once
    const unsigned blocks = (unsigned)((n + 63) / 64), tail_blocks = (tc + 63) / 64;
    if (static_pairs) {
      sample_kernel<true><<<blocks, 64, 0, s>>>(pc, tab, seed, first + (uint64_t)off, work, counters, tail_t, tail_state, tc, rows, row_index, cap, n_out, st);
      sample_tail_kernel<true><<<tail_blocks, 64, 0, s>>>(pc, tab, seed, first + (uint64_t)off, counters, tail_t, tail_state, tc, rows, row_index, cap, n_out, st);
    } else {
      sample_kernel<false><<<blocks, 64, 0, s>>>(pc, tab, seed, first + (uint64_t)off, work, counters, tail_t, tail_state, tc, rows, row_index, cap, n_out, st);
      sample_tail_kernel<false><<<tail_blocks, 64, 0, s>>>(pc, tab, seed, first + (uint64_t)off, counters, tail_t, tail_state, tc, rows, row_index, cap, n_out, st);
    }
  }
  return cudaGetLastError();
}

}  // namespace nwb
