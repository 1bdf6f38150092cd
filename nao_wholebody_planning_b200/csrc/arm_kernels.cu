// arm_kernels.cu - batched right-arm operators for sm_100a:
//   arm_ik           hand position + direction -> all (<= 8) joint solutions     (IKF:341-1517 in closed form)
//   goal_arm_ik      computeRightArmIK, generated-solver branch                   (NCK:224-506)
//   enforce_ma       enforce_MA_Constraints for independent (q_near, q_new) pairs (RP:597-696)
//   goal_sample      sample_stable_configs, goal branch                           (SCG:278-436)
//   ma_project       the same constraint inside the lock-step RRT driver          (RP:829-837)
// IKF = rrt_connect_planner/src/ikfast_TranslationDirection5D.cpp, NCK = .../nao_constraint_kinematics.cpp,
// RP = .../rrt_planner.cpp, SCG = .../stable_config_generator.cpp.
// One unit (target, configuration, sample, query) per thread in fp64; the per-thread building blocks
// are in arm_core.cuh / planner_core.cuh, the validity test in validity_core.cuh (fp32).
#include <cuda_runtime.h>

#include "arm_core.cuh"
#include "nwb_internal.h"
#include "validity_core.cuh"

namespace nwb {

struct Pose6 {
  double v[6];
};

__global__ void __launch_bounds__(128)
arm_ik_kernel(const double* __restrict__ targets /*[n][6]*/, int64_t n, double* __restrict__ sols /*[n][8][5]*/,
              int32_t* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* t = targets + i * 6;
  double* o = sols + i * 40;
  int k = 0;
  int c = arm5d_ik(d3(t[0], t[1], t[2]), d3(t[3], t[4], t[5]), [&](const double* q) {
    for (int j = 0; j < 5; ++j) o[k * 5 + j] = q[j];
    ++k;
  });
  counts[i] = c;
}

__global__ void __launch_bounds__(128)
goal_arm_ik_kernel(const __grid_constant__ PlanConst pc, const __grid_constant__ Pose6 pose, uint64_t seed, uint64_t first_index,
                   const double* __restrict__ q, int64_t n, uint8_t* __restrict__ ok, double* __restrict__ arm,
                   double* __restrict__ erg) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double rl[5];
  for (int m = 0; m < 5; ++m) rl[m] = -q[i * NWB_NJ + m];
  DFrame hip;
  right_leg_fk(rl, hip);
  double a[5] = {0, 0, 0, 0, 0}, e = 0.0;
  ok[i] = goal_right_arm_ik(hip, pose.v, pc, seed, first_index + (uint64_t)i, a, &e);
  for (int m = 0; m < 5; ++m) arm[i * 5 + m] = a[m];
  erg[i] = e;
}

__global__ void __launch_bounds__(128)
enforce_ma_kernel(const __grid_constant__ PlanConst pc, MaView ma, int start_tree, const int32_t* __restrict__ near_hand,
                  const double* __restrict__ q_near, const double* __restrict__ q_ds, int64_t n, double* __restrict__ q_out,
                  int32_t* __restrict__ status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[NWB_NJ], q[NWB_NJ];
  for (int j = 0; j < NWB_NJ; ++j) {
    a[j] = q_near[i * NWB_NJ + j];
    q[j] = q_ds[i * NWB_NJ + j];
  }
  const int st = enforce_ma_constraints(ma, start_tree != 0, near_hand[i], a, q, pc);
  status[i] = st;
  if (st != NWB_TRAPPED)
    for (int j = 0; j < NWB_NJ; ++j) q_out[i * NWB_NJ + j] = q[j];
}

// thread-private point table (as planner_kernels.cu)
struct ArmLocalStore {
  float v[nwbm::kPointFloats];
  __device__ __forceinline__ void put(int slot, float x) { v[slot] = x; }
  __device__ __forceinline__ float get(int slot) const { return v[slot]; }
};

// Goal branch of sample_stable_configs: sample i draws its 21 joints exactly as the database branch
// does (same stream, same slots), then LHipYawPitch = 0 and the left arm is pinned (SCG:291-300).
// The launch covers nq independent requests (hand pose + seed each) x `count` sample indices; accepted
// rows are compacted per request with their sample index and ergonomics index; the caller keeps the
// six lowest indices (the serial loop stops at `goal_samples > 5`, SCG:429).
template <bool STATIC>
__global__ void __launch_bounds__(64)
goal_sample_kernel(const __grid_constant__ PlanConst pc, const __grid_constant__ ValidityTables tab, const double* __restrict__ poses,
                   const uint64_t* __restrict__ seeds, int nq, uint64_t first, int64_t count, int max_ik_iter,
                   double* __restrict__ rows, uint64_t* __restrict__ row_index, double* __restrict__ row_erg, int64_t cap,
                   unsigned long long* __restrict__ n_out, uint8_t* __restrict__ stage_out, double* __restrict__ erg_all,
                   double* __restrict__ rows_all, int32_t* __restrict__ row_req) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * nq) return;
  const int qi = (int)(t / count);
  const uint64_t i = first + (uint64_t)(t - (int64_t)qi * count);
  const uint64_t seed = seeds[qi];
  double pose[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) pose[j] = poses[qi * 6 + j];
  double q[NWB_NJ];
#pragma unroll 1
  for (int blk = 0; blk < (NWB_NJ + 3) / 4; ++blk) {
    uint32_t r[4];
    philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)blk, kStreamSampler, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int l = 0; l < 4; ++l) {
      int j = 4 * blk + l;
      if (j < NWB_NJ) q[j] = __dadd_rn(__dmul_rn(pc.jmax[j] - pc.jmin[j], u01(r[l])), pc.jmin[j]);
    }
  }
  q[15] = 0.0;
  q[5] = 1.5708; q[6] = 0.17; q[7] = 0.0; q[8] = -0.036; q[9] = 0.0;     // SCG:296-300
  double rl[5], ll[5], arm[5], e = 0.0;
  for (int m = 0; m < 5; ++m) rl[m] = -q[m];
  DFrame hip, target;
  right_leg_fk(rl, hip);
  uint8_t stage = 0;
  // The reference solves the left leg first and the right arm second (SCG:318-347); a sample survives
  // only if BOTH succeed and neither consumes random numbers, so the order is free: the closed-form arm
  // solve (cheap, rejects most samples) runs first and only its survivors pay for the NR iteration.
  bool solved = max_ik_iter > 0 && goal_right_arm_ik(hip, pose, pc, seed, i, arm, &e);
  if (solved) {
    desired_left_leg_pose(hip, target);
    for (int m = 0; m < 5; ++m) ll[m] = -rl[m];          // NCK:150 seed; further attempts repeat the same arithmetic (quirk C5)
    solved = left_leg_ik_nr(target, ll, pc.ik_maxiter, pc.ik_eps, pc.pinv_eps, nullptr) && limits_ok(ll, pc.jmin + 16, pc.jmax + 16, 5);
  }
  if (solved) {
    for (int m = 0; m < 5; ++m) {
      q[16 + m] = ll[m];
      q[10 + m] = arm[m];
    }
    float qf[NWB_NJ];
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) qf[j] = (float)q[j];
    ArmLocalStore st;
    ValidityResult<float> r = evaluate_config<float, ArmLocalStore, false, false, STATIC ? PAIRS_STATIC_GROUP : PAIRS_TABLE, /*ENV_CULL=*/true, /*BIG_SCENE=*/true>(tab, qf, st, nullptr, nullptr);
    const bool feasible = !r.hit && r.stable;
    stage = feasible ? 2 : 1;
    if (feasible) {
      // row_req: all requests share ONE pool of cap rows (the host fetches only the rows in use), else cap rows each
      unsigned long long slot = atomicAdd(row_req ? n_out : n_out + qi, 1ull);
      if ((int64_t)slot < cap) {
        const size_t o = row_req ? (size_t)slot : (size_t)qi * cap + slot;
        if (row_req) row_req[o] = qi;
        for (int j = 0; j < NWB_NJ; ++j) rows[o * NWB_NJ + j] = q[j];
        row_index[o] = i;
        row_erg[o] = e;
      }
    }
  } else {
    e = 0.0;
  }
  if (stage_out) stage_out[t] = stage;
  if (erg_all) erg_all[t] = e;
  if (rows_all)
    for (int j = 0; j < NWB_NJ; ++j) rows_all[t * NWB_NJ + j] = q[j];
}

// ---- the same sampler in three phases (service path: no per-sample outputs) ------------------------------------------
// A request keeps ~0.2 % of its samples, and what rejects the others is cheap for most of them and expensive for a
// few: in the one-kernel form every warp waits for its slowest lane (one 100-iteration leg solve stalls 31 rejected
// neighbours).  So: (1) every sample runs the two exact reach tests, (2) the survivors - packed densely - run the
// closed-form arm solve, (3) ITS survivors - packed again - run the leg solve and the validity test.  Each phase
// redraws the sample from its counter (cheaper than carrying 21 doubles through memory).
struct GoalDraw {
  int qi;
  uint64_t i, seed;
  double q[NWB_NJ], rl[5];
  DFrame hip;
};
__device__ __forceinline__ void goal_draw(const PlanConst& pc, const uint64_t* __restrict__ seeds, uint64_t first, int64_t count,
                                          int64_t t, GoalDraw& g) {
  g.qi = (int)(t / count);
  g.i = first + (uint64_t)(t - (int64_t)g.qi * count);
  g.seed = seeds[g.qi];
#pragma unroll 1
  for (int blk = 0; blk < (NWB_NJ + 3) / 4; ++blk) {
    uint32_t r[4];
    philox4x32_10((uint32_t)g.i, (uint32_t)(g.i >> 32), (uint32_t)blk, kStreamSampler, (uint32_t)g.seed, (uint32_t)(g.seed >> 32), r);
    for (int l = 0; l < 4; ++l) {
      int j = 4 * blk + l;
      if (j < NWB_NJ) g.q[j] = __dadd_rn(__dmul_rn(pc.jmax[j] - pc.jmin[j], u01(r[l])), pc.jmin[j]);
    }
  }
  g.q[15] = 0.0;
  g.q[5] = 1.5708; g.q[6] = 0.17; g.q[7] = 0.0; g.q[8] = -0.036; g.q[9] = 0.0;     // SCG:296-300
  for (int m = 0; m < 5; ++m) g.rl[m] = -g.q[m];
  right_leg_fk(g.rl, g.hip);
}
// dense append of the lanes with `keep` to list[]: one atomic per warp; returns the lane's slot (undefined if !keep)
__device__ __forceinline__ unsigned warp_append(bool keep, unsigned* counter) {
  const unsigned mask = __ballot_sync(0xffffffffu, keep);
  if (!mask) return 0;
  const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
}

// counters[0] = |work1|, counters[1] = |work2|
__global__ void __launch_bounds__(128)
goal_filter_kernel(const __grid_constant__ PlanConst pc, const double* __restrict__ poses, const uint64_t* __restrict__ seeds, int nq,
                   uint64_t first, int64_t count, int max_ik_iter, uint32_t* __restrict__ work1, unsigned* __restrict__ counters) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  if (t < count * nq && max_ik_iter > 0) {
    GoalDraw g;
    goal_draw(pc, seeds, first, count, t, g);
    const double* pose = poses + g.qi * 6;
    DFrame target;
    desired_left_leg_pose(g.hip, target);
    keep = left_leg_target_feasible(target, pc.ik_eps);
    if (keep && !(pose[3] == 0.0 && pose[4] == 0.0 && pose[5] == 0.0)) {
      // |grasp point - shoulder| <= |upper arm| + forearm + hand offset (scaled by |direction|): a necessary
      // condition of the closed-form solve, with 1e-6 of slack against rounding
      double p6[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) p6[j] = pose[j];
      D3 pos, dir;
      hand_target_in_torso(g.hip, p6, pos, dir);
      const double dn = sqrt(ddot(dir, dir));
      const double reach = sqrt(kArmUpperX * kArmUpperX + kArmUpperY * kArmUpperY) + kArmForearm + kArmHandOffset * fmax(1.0, dn) + 1e-6;
      const double rx = pos.x, ry = pos.y - kArmShoulderY, rz = pos.z - kArmShoulderZ;
      keep = !(rx * rx + ry * ry + rz * rz > reach * reach);
    }
  }
  const unsigned slot = warp_append(keep, counters);
  if (keep) work1[slot] = (uint32_t)t;
}

__global__ void __launch_bounds__(64)
goal_arm_kernel(const __grid_constant__ PlanConst pc, const double* __restrict__ poses, const uint64_t* __restrict__ seeds,
                uint64_t first, int64_t count, const uint32_t* __restrict__ work1, unsigned* __restrict__ counters,
                uint32_t* __restrict__ work2, double* __restrict__ work2_arm /*[.][6]: 5 joints + ergonomics*/) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  uint32_t t = 0;
  double arm[5], e = 0.0;
  if (w < (int64_t)counters[0]) {
    t = work1[w];
    GoalDraw g;
    goal_draw(pc, seeds, first, count, (int64_t)t, g);
    double pose[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) pose[j] = poses[g.qi * 6 + j];
    keep = goal_right_arm_ik(g.hip, pose, pc, g.seed, g.i, arm, &e);
  }
  const unsigned slot = warp_append(keep, counters + 1);
  if (keep) {
    work2[slot] = t;
#pragma unroll
    for (int m = 0; m < 5; ++m) work2_arm[(size_t)slot * 6 + m] = arm[m];
    work2_arm[(size_t)slot * 6 + 5] = e;
  }
}

template <bool STATIC>
__global__ void __launch_bounds__(64)
goal_solve_kernel(const __grid_constant__ PlanConst pc, const __grid_constant__ ValidityTables tab, const uint64_t* __restrict__ seeds,
                  uint64_t first, int64_t count, const unsigned* __restrict__ counters, const uint32_t* __restrict__ work2,
                  const double* __restrict__ work2_arm, double* __restrict__ rows, uint64_t* __restrict__ row_index,
                  double* __restrict__ row_erg, int64_t cap, unsigned long long* __restrict__ n_out, int32_t* __restrict__ row_req) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= (int64_t)counters[1]) return;
  GoalDraw g;
  goal_draw(pc, seeds, first, count, (int64_t)work2[w], g);
  DFrame target;
  desired_left_leg_pose(g.hip, target);
  double ll[5];
  for (int m = 0; m < 5; ++m) ll[m] = -g.rl[m];          // NCK:150 seed (quirk C5 as above)
  if (!(left_leg_ik_nr(target, ll, pc.ik_maxiter, pc.ik_eps, pc.pinv_eps, nullptr) && limits_ok(ll, pc.jmin + 16, pc.jmax + 16, 5))) return;
  for (int m = 0; m < 5; ++m) {
    g.q[16 + m] = ll[m];
    g.q[10 + m] = work2_arm[(size_t)w * 6 + m];
  }
  float qf[NWB_NJ];
#pragma unroll
  for (int j = 0; j < NWB_NJ; ++j) qf[j] = (float)g.q[j];
  ArmLocalStore st;
  ValidityResult<float> r = evaluate_config<float, ArmLocalStore, false, false, STATIC ? PAIRS_STATIC_GROUP : PAIRS_TABLE, /*ENV_CULL=*/true, /*BIG_SCENE=*/true>(tab, qf, st, nullptr, nullptr);
  if (r.hit || !r.stable) return;
  unsigned long long slot = atomicAdd(row_req ? n_out : n_out + g.qi, 1ull);
  if ((int64_t)slot < cap) {
    const size_t o = row_req ? (size_t)slot : (size_t)g.qi * cap + slot;
    if (row_req) row_req[o] = g.qi;
    for (int j = 0; j < NWB_NJ; ++j) rows[o * NWB_NJ + j] = g.q[j];
    row_index[o] = g.i;
    row_erg[o] = work2_arm[(size_t)w * 6 + 5];
  }
}

// enforce_MA_Constraints inside extendTree (RP:829-837) for the lock-step driver: one thread per query.
__global__ void __launch_bounds__(32)
ma_project_kernel(const __grid_constant__ PlanConst pc, const double* __restrict__ ma_pts, int n_pts, ForestView F, RrtState S,
                  const int32_t* __restrict__ idx_a, const double* __restrict__ q_near, double* __restrict__ q_new,
                  int32_t* __restrict__ status, int nq) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nq || !S.active[k] || status[k] == NWB_TRAPPED) return;
  const int t = S.tree_a[k];
  MaView ma{ma_pts + (size_t)k * n_pts * 6, n_pts};
  double a[NWB_NJ], q[NWB_NJ];
  for (int j = 0; j < NWB_NJ; ++j) {
    a[j] = q_near[(size_t)k * NWB_NJ + j];
    q[j] = q_new[(size_t)k * NWB_NJ + j];
  }
  const int near_hand = F.hand[(size_t)t * F.cap + idx_a[k]];
  const int st = enforce_ma_constraints(ma, (t & 1) == 0, near_hand, a, q, pc);
  if (st == NWB_TRAPPED) {
    status[k] = NWB_TRAPPED;
    return;
  }
  if (st == NWB_ADVANCED) {
    status[k] = NWB_ADVANCED;
    for (int j = 10; j < 15; ++j) q_new[(size_t)k * NWB_NJ + j] = q[j];
  }
}

// ---- launchers ---------------------------------------------------------------------------------------
cudaError_t launch_arm_ik(const double* targets, int64_t n, double* sols, int32_t* counts, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  arm_ik_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(targets, n, sols, counts);
  return cudaGetLastError();
}
cudaError_t launch_goal_arm_ik(const PlanConst& pc, const double* pose6, uint64_t seed, uint64_t first_index, const double* q,
                               int64_t n, uint8_t* ok, double* arm, double* erg, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  Pose6 p;
  for (int j = 0; j < 6; ++j) p.v[j] = pose6[j];
  goal_arm_ik_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(pc, p, seed, first_index, q, n, ok, arm, erg);
  return cudaGetLastError();
}
cudaError_t launch_enforce_ma(const PlanConst& pc, const double* pts_dev, int n_pts, int start_tree, const int32_t* near_hand,
                              const double* q_near, const double* q_ds, int64_t n, double* q_out, int32_t* status, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  enforce_ma_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(pc, MaView{pts_dev, n_pts}, start_tree, near_hand, q_near, q_ds, n,
                                                               q_out, status);
  return cudaGetLastError();
}
cudaError_t launch_goal_sample(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, const double* poses_dev,
                               const uint64_t* seeds_dev, int nq, uint64_t first, int64_t count, int max_ik_iter, double* rows,
                               uint64_t* row_index, double* row_erg, int64_t cap, unsigned long long* n_out, uint8_t* stage_out,
                               double* erg_all, double* rows_all, int32_t* row_req, cudaStream_t s) {
  if (count <= 0 || nq <= 0) return cudaSuccess;
  unsigned blocks = (unsigned)((count * nq + 63) / 64);
  if (static_pairs)
    goal_sample_kernel<true><<<blocks, 64, 0, s>>>(pc, tab, poses_dev, seeds_dev, nq, first, count, max_ik_iter, rows, row_index,
                                                   row_erg, cap, n_out, stage_out, erg_all, rows_all, row_req);
  else
    goal_sample_kernel<false><<<blocks, 64, 0, s>>>(pc, tab, poses_dev, seeds_dev, nq, first, count, max_ik_iter, rows, row_index,
                                                    row_erg, cap, n_out, stage_out, erg_all, rows_all, row_req);
  return cudaGetLastError();
}
size_t goal_sample_phased_scratch_bytes(int64_t items) { return 256 + (size_t)items * (4 + 4 + 48); }
cudaError_t launch_goal_sample_phased(const PlanConst& pc, const ValidityTables& tab, bool static_pairs, const double* poses_dev,
                                      const uint64_t* seeds_dev, int nq, uint64_t first, int64_t count, int max_ik_iter, double* rows,
                                      uint64_t* row_index, double* row_erg, int64_t cap, unsigned long long* n_out, int32_t* row_req,
                                      void* scratch, cudaStream_t s) {
  if (count <= 0 || nq <= 0) return cudaSuccess;
  const int64_t items = count * nq;
  unsigned* counters = static_cast<unsigned*>(scratch);
  double* work2_arm = reinterpret_cast<double*>(static_cast<char*>(scratch) + 256);
  uint32_t* work1 = reinterpret_cast<uint32_t*>(work2_arm + (size_t)items * 6);
  uint32_t* work2 = work1 + items;
  cudaError_t e = cudaMemsetAsync(counters, 0, 8, s);
  if (e != cudaSuccess) return e;
  // the list sizes stay on the device: phases 2 and 3 are launched for the worst case and surplus blocks leave at once
  goal_filter_kernel<<<(unsigned)((items + 127) / 128), 128, 0, s>>>(pc, poses_dev, seeds_dev, nq, first, count, max_ik_iter, work1, counters);
  const unsigned blocks = (unsigned)((items + 63) / 64);
  goal_arm_kernel<<<blocks, 64, 0, s>>>(pc, poses_dev, seeds_dev, first, count, work1, counters, work2, work2_arm);
  if (static_pairs)
    goal_solve_kernel<true><<<blocks, 64, 0, s>>>(pc, tab, seeds_dev, first, count, counters, work2, work2_arm, rows, row_index, row_erg, cap, n_out, row_req);
  else
    goal_solve_kernel<false><<<blocks, 64, 0, s>>>(pc, tab, seeds_dev, first, count, counters, work2, work2_arm, rows, row_index, row_erg, cap, n_out, row_req);
  return cudaGetLastError();
}
cudaError_t launch_ma_project(const PlanConst& pc, const double* ma_pts, int n_pts, ForestView F, RrtState S, const int32_t* idx_a,
                              const double* q_near, double* q_new, int32_t* status, int nq, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  ma_project_kernel<<<(unsigned)((nq + 31) / 32), 32, 0, s>>>(pc, ma_pts, n_pts, F, S, idx_a, q_near, q_new, status, nq);
  return cudaGetLastError();
}

}  // namespace nwb
