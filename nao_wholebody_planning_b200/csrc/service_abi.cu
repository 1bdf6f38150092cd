// service_abi.cu - extern "C" entry points of the right-arm operators, goal-configuration sampling and
// the planning services (include/nwb.h).  Host-side orchestration only; the arithmetic runs in
// arm_kernels.cu / planner_kernels.cu / validity.cu.
//
// Reference code replaced (file:line relative to the reference checkout):
//   NAO_PLANNER_CONTROL::compute_goal_config / compute_motion_plan / compute_linear_manipulation_plan /
//   compute_circular_manipulation_plan        rrt_connect_planner/src/nao_planner_control.cpp:150-870
//   StableConfigGenerator::sample_stable_configs (goal branch), sort_goal_configs_by_ergonomy, loadGoalConfig
//                                              rrt_connect_planner/src/stable_config_generator.cpp:176-507,510-661,666-756
//   ArticulatedObjectConstraints::computeDrawerTrajectoryPoints / computeDoorTrajectoryPoints
//                                              rrt_connect_planner/src/articulated_object_constraints.cpp:105-189
//   RRT_Planner::activate_drawer_constraint / activate_door_constraint / writePath / generateTrajectory (time stamps)
//                                              rrt_connect_planner/src/rrt_planner.cpp:195-223,1303-1455,1532-1553
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "nwb_internal.h"
#include "planner_core.cuh"

using namespace nwb;

namespace nwb {
void make_plan_const(const nwb_ctx* ctx, const double* weights, double step, PlanConst& pc);
}

namespace {
#define S_CUDA(ctx, call) NWB_TRY_CUDA(ctx, call)

int bad(nwb_ctx* ctx, const char* msg) {
  if (ctx) ctx->err = msg;
  return NWB_E_INVALID;
}
bool static_pairs(const nwb_ctx* ctx) { return ctx->params.srdf_trim != 0 && !ctx->force_table_pairs; }

using Scratch = nwb::DeviceArena;

// the goal files keep 6 significant digits: `ofstream << double` (SCG:409-417), re-read with strtod (SCG:540-560)
double through_text(double v) {
  char buf[64];
  std::snprintf(buf, sizeof buf, "%g", v);
  return std::strtod(buf, nullptr);
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct GoalSet {
  int n_found = 0;
  int64_t drawn = 0;
  double rows[6][NWB_NJ];
  double erg[6];
  uint64_t index[6];
};

// Host-side finishing of a BATCH (text round trip of the goal rows, time parameterisation of the trajectories) is
// independent per request: spread it over the host cores.  NWB_HOST_THREADS overrides the thread count.
template <class F>
void parallel_for(int n, F&& body) {
  int threads = (int)std::thread::hardware_concurrency();
  if (const char* e = std::getenv("NWB_HOST_THREADS")) threads = std::atoi(e);
  threads = std::max(1, std::min({threads, 32, n}));
  if (threads == 1) {
    for (int i = 0; i < n; ++i) body(i);
    return;
  }
  std::atomic<int> next{0};
  auto worker = [&] {
    for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) body(i);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
}
bool trace_on() { return std::getenv("NWB_TRACE") != nullptr; }

constexpr int kGoalsWanted = 6;   // SCG:429: the loop leaves once goal_samples > 5

// first six accepted rows (sample order) -> text round trip -> sort_goal_configs_by_ergonomy
void finish_goal_set(const std::vector<double>& rows, const std::vector<double>& erg, const std::vector<uint64_t>& idx, int max_samples,
                     GoalSet& g) {
  const size_t got = idx.size();
  std::vector<size_t> order(got);
  std::iota(order.begin(), order.end(), (size_t)0);
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return idx[a] < idx[b]; });
  const int n = (int)std::min<size_t>(got, kGoalsWanted);
  g.drawn = n == kGoalsWanted ? (int64_t)idx[order[n - 1]] + 1 : max_samples;
  // file order = sample order; values pass through the text format before the sort reads them
  double f_erg[6], f_rows[6][NWB_NJ];
  uint64_t f_idx[6];
  for (int r = 0; r < n; ++r) {
    f_erg[r] = through_text(erg[order[r]]);
    f_idx[r] = idx[order[r]];
    for (int j = 0; j < NWB_NJ; ++j) f_rows[r][j] = through_text(rows[order[r] * NWB_NJ + j]);
  }
  // SCG:585-612: n rounds of "largest remaining value > 0.0"; the winner is zeroed afterwards
  double e2[6];
  for (int r = 0; r < n; ++r) e2[r] = f_erg[r];
  int imax = 0;
  for (int round = 0; round < n; ++round) {
    double vmax = 0.0;
    for (int i = 0; i < n; ++i)
      if (e2[i] > vmax) {
        vmax = e2[i];
        imax = i;
      }
    g.erg[round] = f_erg[imax];
    g.index[round] = f_idx[imax];
    std::memcpy(g.rows[round], f_rows[imax], sizeof(double) * NWB_NJ);
    e2[imax] = 0.0;
  }
  g.n_found = n;
}

// sample_stable_configs(goal branch) + sort_goal_configs_by_ergonomy for nq independent requests.
// All nq x max_samples candidates are evaluated in ONE launch (the serial reference stops after the
// sixth accepted sample; the samples behind it are independent, so evaluating them costs idle lanes,
// not time) and per request the six accepted rows with the lowest sample index are kept.
int sample_goals_batch(nwb_ctx* ctx, int nq, const uint64_t* seeds, const double* poses, int max_samples, int max_ik_iter,
                       GoalSet* out) {
  for (int k = 0; k < nq; ++k) out[k] = GoalSet();
  if (max_samples <= 0 || nq <= 0) return NWB_OK;
  S_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  // The accepted samples of all requests go into ONE pool (a request accepts ~0.3 % of its samples, so 512 rows
  // per request on average are plenty) and only the rows in use come back.  A batch that outgrows the pool is
  // redone one request at a time, each with room for every sample.
  const int64_t cap = nq == 1 ? max_samples : (int64_t)nq * std::min<int64_t>(max_samples, 512);
  Scratch A(ctx);
  size_t o_rows = A.add(sizeof(double) * NWB_NJ * cap), o_idx = A.add(sizeof(uint64_t) * cap), o_erg = A.add(sizeof(double) * cap),
         o_req = A.add(sizeof(int32_t) * cap), o_n = A.add(8), o_pose = A.add(sizeof(double) * 6 * nq),
         o_seed = A.add(sizeof(uint64_t) * nq), o_work = A.add(goal_sample_phased_scratch_bytes((int64_t)nq * max_samples));
  int rc = A.commit();
  if (rc != NWB_OK) return rc;
  PlanConst pc;
  make_plan_const(ctx, nullptr, ctx->params.planner_step_factor, pc);
  S_CUDA(ctx, cudaMemsetAsync(A.at<char>(o_n), 0, 8, s));
  S_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_pose), poses, sizeof(double) * 6 * nq, cudaMemcpyHostToDevice, s));
  S_CUDA(ctx, cudaMemcpyAsync(A.at<uint64_t>(o_seed), seeds, sizeof(uint64_t) * nq, cudaMemcpyHostToDevice, s));
  if (std::getenv("NWB_GOAL_ONE_KERNEL"))                   // the one-launch form of the same sampler (cross-check / timing)
    S_CUDA(ctx, launch_goal_sample(pc, ctx->tab_group, static_pairs(ctx), A.at<double>(o_pose), A.at<uint64_t>(o_seed), nq, 0, max_samples,
                                   max_ik_iter, A.at<double>(o_rows), A.at<uint64_t>(o_idx), A.at<double>(o_erg), cap,
                                   A.at<unsigned long long>(o_n), nullptr, nullptr, nullptr, A.at<int32_t>(o_req), s));
  else
    S_CUDA(ctx, launch_goal_sample_phased(pc, ctx->tab_group, static_pairs(ctx), A.at<double>(o_pose), A.at<uint64_t>(o_seed), nq, 0,
                                          max_samples, max_ik_iter, A.at<double>(o_rows), A.at<uint64_t>(o_idx), A.at<double>(o_erg), cap,
                                          A.at<unsigned long long>(o_n), A.at<int32_t>(o_req), A.at<char>(o_work), s));
  unsigned long long got = 0;
  const double t_launch = now_s();
  S_CUDA(ctx, cudaMemcpyAsync(&got, A.at<char>(o_n), 8, cudaMemcpyDeviceToHost, s));
  S_CUDA(ctx, cudaStreamSynchronize(s));
  const double t_kernel = now_s();
  ctx->last_launches = std::getenv("NWB_GOAL_ONE_KERNEL") ? 1 : 3;
  if ((int64_t)got > cap) {                                  // nq > 1 only: a single request's pool holds every sample
    for (int k = 0; k < nq; ++k) {
      rc = sample_goals_batch(ctx, 1, seeds + k, poses + 6 * k, max_samples, max_ik_iter, out + k);
      if (rc != NWB_OK) return rc;
    }
    return NWB_OK;
  }
  const size_t n = (size_t)got;
  std::vector<double> rows(n * NWB_NJ), erg(n);
  std::vector<uint64_t> idx(n);
  std::vector<int32_t> req(n);
  if (n) {
    S_CUDA(ctx, cudaMemcpyAsync(rows.data(), A.at<double>(o_rows), sizeof(double) * NWB_NJ * n, cudaMemcpyDeviceToHost, s));
    S_CUDA(ctx, cudaMemcpyAsync(idx.data(), A.at<uint64_t>(o_idx), sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, s));
    S_CUDA(ctx, cudaMemcpyAsync(erg.data(), A.at<double>(o_erg), sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    S_CUDA(ctx, cudaMemcpyAsync(req.data(), A.at<int32_t>(o_req), sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    S_CUDA(ctx, cudaStreamSynchronize(s));
  }
  std::vector<std::vector<double>> k_rows(nq), k_erg(nq);
  std::vector<std::vector<uint64_t>> k_idx(nq);
  for (size_t r = 0; r < n; ++r) {
    const int k = req[r];
    k_rows[k].insert(k_rows[k].end(), rows.begin() + r * NWB_NJ, rows.begin() + (r + 1) * NWB_NJ);
    k_erg[k].push_back(erg[r]);
    k_idx[k].push_back(idx[r]);
  }
  const double t_fetch = now_s();
  parallel_for(nq, [&](int k) { finish_goal_set(k_rows[k], k_erg[k], k_idx[k], max_samples, out[k]); });
  if (trace_on())
    std::fprintf(stderr, "[nwb] goal sampling nq=%d accepted=%llu: kernel %.3f ms, fetch %.3f ms, finish %.3f ms\n", nq, got,
                 (t_kernel - t_launch) * 1e3, (t_fetch - t_kernel) * 1e3, (now_s() - t_fetch) * 1e3);
  return NWB_OK;
}

int sample_goals(nwb_ctx* ctx, uint64_t seed, const double* pose6, int max_samples, int max_ik_iter, GoalSet& g) {
  return sample_goals_batch(ctx, 1, &seed, pose6, max_samples, max_ik_iter, &g);
}

void limb_weights(double r_arm, double l_arm, double legs, double* w) {   // NPC:280-306
  for (int i = 0; i < 5; ++i) w[i] = legs;
  for (int i = 5; i < 10; ++i) w[i] = l_arm;
  for (int i = 10; i < 15; ++i) w[i] = r_arm;
  w[15] = 0.0;                                                            // LHipYawPitch
  for (int i = 16; i < NWB_NJ; ++i) w[i] = legs;
}

// ---- trajectory_processing::IterativeParabolicTimeParameterization (MoveIt, called RP:1610-1611) -----------------
// Restated from the published algorithm (moveit_core/trajectory_processing, Hydro/Indigo vintage): MoveIt is
// not vendored in the reference.  time_diff[i] = duration of the interval between way-points i and i+1.
namespace iptp {
constexpr double kRoundingThreshold = 0.01;

void apply_velocity_constraints(const double* q, int n, const double* v_max, std::vector<double>& time_diff) {
  for (int i = 0; i < n - 1; ++i)
    for (int j = 0; j < NWB_NJ; ++j) {
      const double dq1 = q[(size_t)i * NWB_NJ + j], dq2 = q[(size_t)(i + 1) * NWB_NJ + j];
      const double t_min = std::fabs(dq2 - dq1) / v_max[j];
      if (t_min > time_diff[i]) time_diff[i] = t_min;
    }
}
// grow dt1 (resp. dt2) by 1 % until the blend acceleration is within a_max
double find_t1(double dq1, double dq2, double dt1, double dt2, double a_max) {
  const double mult_factor = 1.01;
  double v1 = dq1 / dt1, v2 = dq2 / dt2, a = 2.0 * (v2 - v1) / (dt1 + dt2);
  while (std::fabs(a) > a_max) {
    v1 = dq1 / dt1;
    v2 = dq2 / dt2;
    a = 2.0 * (v2 - v1) / (dt1 + dt2);
    dt1 *= mult_factor;
  }
  return dt1;
}
double find_t2(double dq1, double dq2, double dt1, double dt2, double a_max) {
  const double mult_factor = 1.01;
  double v1 = dq1 / dt1, v2 = dq2 / dt2, a = 2.0 * (v2 - v1) / (dt1 + dt2);
  while (std::fabs(a) > a_max) {
    v1 = dq1 / dt1;
    v2 = dq2 / dt2;
    a = 2.0 * (v2 - v1) / (dt1 + dt2);
    dt2 *= mult_factor;
  }
  return dt2;
}
void apply_acceleration_constraints(const double* q, int n, const double* a_lim, int max_iterations, double max_time_change,
                                    std::vector<double>& time_diff) {
  int num_updates = 0, iteration = 0;
  bool backwards = false;
  do {
    num_updates = 0;
    ++iteration;
    for (int j = 0; j < NWB_NJ; ++j) {                   // joints outermost: an enlarged interval propagates along the path
      for (int count = 0; count < 2; ++count) {          // forwards, then backwards
        for (int i = 0; i < n - 1; ++i) {
          const int index = backwards ? (n - 1) - i : i;
          const double a_max = a_lim[j];
          const double cur = q[(size_t)index * NWB_NJ + j];
          double q1, q2, q3, dt1, dt2;
          if (index == 0) {                              // first point: mirrored neighbour
            q1 = q[(size_t)(index + 1) * NWB_NJ + j]; q2 = cur; q3 = q1;
            dt1 = dt2 = time_diff[index];
          } else if (index < n - 1) {
            q1 = q[(size_t)(index - 1) * NWB_NJ + j]; q2 = cur; q3 = q[(size_t)(index + 1) * NWB_NJ + j];
            dt1 = time_diff[index - 1];
            dt2 = time_diff[index];
          } else {                                       // last point: only n-1 intervals
            q1 = q[(size_t)(index - 1) * NWB_NJ + j]; q2 = cur; q3 = q1;
            dt1 = dt2 = time_diff[index - 1];
          }
          double a = 0.0;
          if (!(dt1 == 0.0 || dt2 == 0.0)) {
            const double v1 = (q2 - q1) / dt1, v2 = (q3 - q2) / dt2;
            a = 2.0 * (v2 - v1) / (dt1 + dt2);
          }
          if (std::fabs(a) > a_max + kRoundingThreshold) {
            if (!backwards) {
              dt2 = std::min(dt2 + max_time_change, find_t2(q2 - q1, q3 - q2, dt1, dt2, a_max));
              time_diff[index] = dt2;
            } else {
              dt1 = std::min(dt1 + max_time_change, find_t1(q2 - q1, q3 - q2, dt1, dt2, a_max));
              time_diff[index - 1] = dt1;
            }
            ++num_updates;
          }
        }
        backwards = !backwards;
      }
    }
  } while (num_updates > 0 && iteration < max_iterations);
}
// computeTimeStamps + the 2 s head start of RP:1614-1618
void time_stamps(const nwb_params& P, const double* q, int n, double* tfs) {
  if (n <= 0) return;
  std::vector<double> time_diff((size_t)std::max(n - 1, 0), 0.0);
  apply_velocity_constraints(q, n, P.joint_max_velocity, time_diff);
  apply_acceleration_constraints(q, n, P.joint_max_acceleration, P.iptp_max_iterations, P.iptp_max_time_change, time_diff);
  double t = 0.0;
  tfs[0] = t + 2.0;
  for (int i = 1; i < n; ++i) {
    t += time_diff[i - 1];
    tfs[i] = t + 2.0;
  }
}
}  // namespace iptp

void put_trajectory(const nwb_params& P, const std::vector<double>& path, bool articulated, bool execute, float time_step_manip,
                    nwb_trajectory& t) {
  const int n = (int)(path.size() / NWB_NJ);
  t.n_points = n;
  float increment = 2.0f;                                                 // RP:1532
  for (int i = 0; i < n; ++i) {
    if (i < t.capacity) {
      if (t.positions) std::memcpy(t.positions + (size_t)i * NWB_NJ, path.data() + (size_t)i * NWB_NJ, sizeof(double) * NWB_NJ);
      if (t.time_from_start) t.time_from_start[i] = (double)increment;
    }
    increment = articulated ? increment + time_step_manip : increment + 1.0f;   // RP:1547-1552
  }
  // RP:1569-1618: everything but an executed articulated plan goes through the time parameterisation
  if (P.time_parameterization && !(execute && articulated) && n > 0 && t.time_from_start) {
    std::vector<double> tfs(n);
    iptp::time_stamps(P, path.data(), n, tfs.data());
    for (int i = 0; i < std::min(n, (int)t.capacity); ++i) t.time_from_start[i] = tfs[i];
  }
}

struct StageOut {
  bool success = false;
  uint32_t num_nodes = 0;
  double search_time = 0.0;
  std::vector<double> path;
};

// setStartGoalConfigs + solveQuery + writePath (RP:227-302, 1039-1212, 1303-1455) for one query
int plan_stage(nwb_ctx* ctx, const double* start, const double* goal, const double* weights, uint64_t seed, const double* hand_traj,
               int n_pts, bool execute, StageOut& out) {
  out = StageOut();
  if (ctx->ds_db.size() < 2 * NWB_NJ) {
    ctx->err = "no stable-configuration database loaded (nwb_ctx_set_ds_database / nwb_ctx_load_ds_database)";
    return NWB_E_STATE;
  }
  const nwb_params& P = ctx->params;
  const int cap = 8192;
  if (ctx->h_raw.size() < (size_t)cap * NWB_NJ) ctx->h_raw.resize((size_t)cap * NWB_NJ);
  std::vector<double>& raw = ctx->h_raw;
  nwb_plan_result r;
  const double t0 = now_s();
  int rc = nwb_solve_query_constrained_batch(ctx, start, goal, 1, ctx->ds_db.data(), (int32_t)(ctx->ds_db.size() / NWB_NJ), weights,
                                             &seed, P.planner_max_iterations, P.planner_step_factor, 0, hand_traj, n_pts, &r,
                                             raw.data(), cap);
  if (rc == NWB_OK && r.success && r.raw_path_len > cap) {                // a raw path longer than the buffer: never truncate,
    raw.resize((size_t)r.raw_path_len * NWB_NJ);                          // fetch it whole (same seed, same search)
    rc = nwb_solve_query_constrained_batch(ctx, start, goal, 1, ctx->ds_db.data(), (int32_t)(ctx->ds_db.size() / NWB_NJ), weights,
                                           &seed, P.planner_max_iterations, P.planner_step_factor, 0, hand_traj, n_pts, &r,
                                           raw.data(), r.raw_path_len);
  }
  if (rc == NWB_E_NOMEM) rc = NWB_OK;       // the search outgrew the largest forest: reported as "no path" (r.success == 0)
  if (rc != NWB_OK) return rc;
  out.search_time = now_s() - t0;                                         // RP:1092-1094: stops before writePath
  if (!r.success) return NWB_OK;
  out.num_nodes = (uint32_t)r.num_nodes_generated;
  const int n_raw = r.raw_path_len;
  if (!hand_traj) {                                                       // RP:1400-1405
    std::vector<double> sel((size_t)(n_raw + 8) * NWB_NJ);
    int n_sel = nwb_shortcut_path(ctx, raw.data(), n_raw, sel.data(), n_raw + 8, nullptr);
    if (n_sel == -100) return NWB_OK;                                     // RP:1919-1923: shortcutting failed, empty trajectory
    if (n_sel < 0) return n_sel;
    if (!execute) {                                                       // RP:1935-1945
      const int rows = (n_sel - 1) * (P.shortcut_waypoints + 1) + 1;
      out.path.resize((size_t)rows * NWB_NJ);
      nwb_interpolate_path(sel.data(), n_sel, P.shortcut_waypoints, out.path.data(), rows);
    } else {
      out.path.assign(sel.begin(), sel.begin() + (size_t)n_sel * NWB_NJ);
    }
  } else if (!execute) {                                                  // RP:1408-1413
    const int rows = (n_raw - 1) * (P.manipulate_waypoints + 1) + 1;
    out.path.resize((size_t)rows * NWB_NJ);
    nwb_interpolate_path(raw.data(), n_raw, P.manipulate_waypoints, out.path.data(), rows);
  } else {
    out.path.assign(raw.begin(), raw.begin() + (size_t)n_raw * NWB_NJ);   // RP:1417
  }
  out.success = !out.path.empty();
  return NWB_OK;
}

// path_shortcutter (RP:1830-1951) for several raw paths at once: every round gathers the candidate edges
// of ALL unfinished paths into one batched edge check, then each path takes its first valid candidate in
// the reference's scan order.  sel[k] = selected way-point indices; ok[k] = false if its 500-round budget ran out.
int shortcut_paths_batch(nwb_ctx* ctx, const std::vector<std::vector<double>>& raw, std::vector<std::vector<int>>& sel,
                         std::vector<uint8_t>& ok) {
  const int n = (int)raw.size(), K = ctx->params.shortcut_waypoints;
  sel.assign(n, {});
  ok.assign(n, 1);
  std::vector<int> cur(n, 0), budget(n, ctx->params.shortcut_max_rounds), W(n);
  std::vector<uint8_t> reached(n, 0);
  for (int k = 0; k < n; ++k) {
    W[k] = (int)(raw[k].size() / NWB_NJ);
    if (W[k] < 1) { reached[k] = 1; ok[k] = 0; continue; }
    sel[k].push_back(0);
  }
  std::vector<double> qa, qb;
  std::vector<uint8_t> valid;
  std::vector<int> cand, first(n + 1);
  for (;;) {
    qa.clear(); qb.clear(); cand.clear();
    bool any = false;
    for (int k = 0; k < n; ++k) {
      first[k] = (int)cand.size();
      if (reached[k] || budget[k] <= 0) continue;
      any = true;
      auto push = [&](int i) {
        cand.push_back(i);
        qa.insert(qa.end(), raw[k].begin() + (size_t)cur[k] * NWB_NJ, raw[k].begin() + (size_t)(cur[k] + 1) * NWB_NJ);
        qb.insert(qb.end(), raw[k].begin() + (size_t)i * NWB_NJ, raw[k].begin() + (size_t)(i + 1) * NWB_NJ);
      };
      push(W[k] - 1);                                       // goal first (RP:1869-1876)
      for (int i = W[k] - 2; i >= 0; --i) {                 // then i = W-2 .. 0 (RP:1884-1916)
        if (i == cur[k]) break;
        push(i);
      }
    }
    first[n] = (int)cand.size();
    if (!any) break;
    valid.resize(cand.size());
    int rc = nwb_is_path_valid_batch(ctx, qa.data(), qb.data(), (int64_t)cand.size(), K, valid.data(), nullptr);
    if (rc != NWB_OK) return rc;
    for (int k = 0; k < n; ++k) {
      if (reached[k] || budget[k] <= 0) continue;
      int pick = -1;
      for (int c = first[k]; c < first[k + 1]; ++c)
        if (valid[c]) { pick = c; break; }
      if (pick == first[k]) {
        sel[k].push_back(W[k] - 1);
        reached[k] = 1;
      } else if (pick >= 0) {
        cur[k] = cand[pick];
        sel[k].push_back(cur[k]);
      } else if (cur[k] + 1 < W[k]) {                       // stuck: advance one raw way-point (RP:1891-1897)
        cur[k] = cur[k] + 1;
        sel[k].push_back(cur[k]);
      }                                                     // else nothing is pushed and the budget runs down (as RP:1884-1916)
      if (--budget[k] == 0) ok[k] = 0;                      // RP:1919-1923 (checked after the loop, reached or not)
    }
  }
  return NWB_OK;
}

void pose_of(const nwb_point3& p, const nwb_point3& d, double* pose6) {
  pose6[0] = p.x; pose6[1] = p.y; pose6[2] = p.z;
  pose6[3] = d.x; pose6[4] = d.y; pose6[5] = d.z;
}

void drawer_points(const double* a, const double* b, int num_interp, double* out) {          // AOC:105-122
  double step[6];
  for (int i = 0; i < 6; ++i) step[i] = (b[i] - a[i]) / (num_interp + 1);
  for (int s = 0; s < num_interp + 2; ++s)
    for (int i = 0; i < 6; ++i) out[s * 6 + i] = a[i] + (step[i] * s);
}
void door_points(const double* a, const double* b, int num_interp, const double* door, double* out) {   // AOC:126-189
  const double relative_angle = door[2] * (M_PI / 180);
  const double step_angle = (door[1] * (M_PI / 180)) / (num_interp + 1);
  double sd[3];
  for (int i = 0; i < 3; ++i) sd[i] = (b[3 + i] - a[3 + i]) / (num_interp + 1);
  double x_door = a[0] + door[3] * std::cos(relative_angle);
  double y_door = a[1] - door[3] * std::sin(relative_angle);
  double z_door = a[2];
  const double x_hinge = x_door - door[0] * std::sin(relative_angle);
  const double y_hinge = y_door - door[0] * std::cos(relative_angle);
  const double z_hinge = z_door;
  for (int s = 0; s < num_interp + 2; ++s) {
    double tmp_angle = ((step_angle * s) - relative_angle);
    x_door = x_hinge - door[0] * std::sin(tmp_angle);
    y_door = y_hinge + door[0] * std::cos(tmp_angle);
    z_door = z_hinge;
    tmp_angle = 1.5708 - tmp_angle;                                       // AOC:176: "90 degrees"
    double* o = out + s * 6;
    o[0] = x_door - door[3] * std::sin(tmp_angle);
    o[1] = y_door - door[3] * std::cos(tmp_angle);
    o[2] = z_door;
    for (int i = 0; i < 3; ++i) o[3 + i] = a[3 + i] + (sd[i] * s);
  }
}

// the common body of the linear / circular manipulation services
int manipulation_plan(nwb_ctx* ctx, const double* pose6, const double* start, bool execute, bool circular, const double* object,
                      nwb_manipulation_plan_response* resp, uint64_t seed) {
  const nwb_params& P = ctx->params;
  resp->motion_plan_approach.n_points = 0;
  resp->motion_plan_manipulation.n_points = 0;
  resp->time_goal_config_approach = resp->time_goal_config_manipulate = 0.0;
  resp->num_nodes_generated_approach = resp->num_nodes_generated_manipulate = 0;
  resp->search_time_approach = resp->search_time_manipulate = 0.0;
  resp->success_approach = resp->success_manipulation = 0;
  // Both hand poses are known from the request alone, so the two goal samplings (NPC:430-436 and 509-516 /
  // 690-696 and 793-800) share ONE launch; the second set is only used if the approach plan succeeds.
  double pm[6];
  const int NI = P.object_interp_points;
  std::vector<double> traj((size_t)(NI + 2) * 6);
  float time_step;
  if (!circular) {
    const double pos_angle = (90.0 - object[0]) * (M_PI / 180);                       // NPC:502
    pm[0] = pose6[0] - object[1] * std::sin(pos_angle);
    pm[1] = pose6[1] + object[1] * std::cos(pos_angle);
    pm[2] = pose6[2];
  } else {                                                                             // NPC:762-790
    const double rel = object[0] * (M_PI / 180);
    double x_door = pose6[0] + object[3] * std::cos(rel);
    double y_door = pose6[1] - object[3] * std::sin(rel);
    const double z_door = pose6[2];
    const double x_hinge = x_door - object[1] * std::sin(rel);
    const double y_hinge = y_door - object[1] * std::cos(rel);
    double tmp_angle = (object[2] - object[0]) * (M_PI / 180);
    x_door = x_hinge - object[1] * std::sin(tmp_angle);
    y_door = y_hinge + object[1] * std::cos(tmp_angle);
    tmp_angle = 1.5708 - tmp_angle;
    pm[0] = x_door - object[3] * std::sin(tmp_angle);
    pm[1] = y_door - object[3] * std::cos(tmp_angle);
    pm[2] = z_door;
  }
  for (int i = 3; i < 6; ++i) pm[i] = pose6[i];
  GoalSet goals[2];
  double poses[12];
  std::memcpy(poses, pose6, sizeof(double) * 6);
  std::memcpy(poses + 6, pm, sizeof(double) * 6);
  const uint64_t gseeds[2] = {seed, seed + 1};
  double t0 = now_s();
  int rc = sample_goals_batch(ctx, 2, gseeds, poses, P.scg_max_samples, P.scg_max_ik_iterations, goals);
  if (rc != NWB_OK) return rc;
  const GoalSet &ga = goals[0], &gm = goals[1];
  resp->time_goal_config_approach = now_s() - t0;        // both sets; time_goal_config_manipulate stays 0
  // ---- approach (NPC:430-494 / 690-753)
  if (ga.n_found == 0) return NWB_OK;
  std::memcpy(resp->wb_goal_config_approach, ga.rows[0], sizeof(double) * NWB_NJ);
  double w[NWB_NJ];
  limb_weights(P.r_arm_weight_approach, P.l_arm_weight_approach, P.legs_weight_approach, w);
  StageOut sa;
  rc = plan_stage(ctx, start, ga.rows[0], w, seed + 2, nullptr, 0, execute, sa);
  if (rc != NWB_OK) return rc;
  resp->num_nodes_generated_approach = sa.num_nodes;
  resp->search_time_approach = sa.search_time;
  resp->success_approach = sa.success;
  if (!sa.success) return NWB_OK;
  put_trajectory(P, sa.path, false, execute, 0.f, resp->motion_plan_approach);
  // ---- manipulation (NPC:498-585 / 757-866)
  if (gm.n_found == 0) return NWB_OK;
  std::memcpy(resp->wb_goal_config_manipulate, gm.rows[0], sizeof(double) * NWB_NJ);
  const float velocity = (float)P.manipulation_velocity;
  if (!circular) {                                                                     // RP:195-206
    drawer_points(pose6, pm, NI, traj.data());
    const float manipulation_time = (float)(object[1]) / velocity;
    time_step = manipulation_time / float(NI + 2);
  } else {                                                                             // NPC:816-822, RP:209-223
    const double door[4] = {object[1], object[2], object[0], object[3]};
    door_points(pose6, pm, NI, door, traj.data());
    const float arc_length = float(door[0] * M_PI * door[1]) / 180.0f;
    const float manipulation_time = arc_length / velocity;
    time_step = manipulation_time / float(NI + 2);
  }
  limb_weights(P.r_arm_weight_manipulate, P.l_arm_weight_manipulate, P.legs_weight_manipulate, w);
  StageOut sm;
  rc = plan_stage(ctx, ga.rows[0], gm.rows[0], w, seed + 3, traj.data(), NI + 2, execute, sm);
  if (rc != NWB_OK) return rc;
  resp->num_nodes_generated_manipulate = sm.num_nodes;
  resp->search_time_manipulate = sm.search_time;
  resp->success_manipulation = sm.success;
  if (sm.success) put_trajectory(P, sm.path, true, execute, time_step, resp->motion_plan_manipulation);
  return NWB_OK;
}
}  // namespace

extern "C" {

int nwb_arm_ik_batch(nwb_ctx* ctx, const double* targets, int64_t n, double* solutions, int32_t* counts) {
  if (!ctx) return NWB_E_INVALID;
  if (n < 0 || (n > 0 && (!targets || !solutions || !counts))) return bad(ctx, "nwb_arm_ik_batch: bad argument");
  if (n == 0) return NWB_OK;
  S_CUDA(ctx, cudaSetDevice(ctx->device));
  Scratch A(ctx);
  size_t o_t = A.add(sizeof(double) * 6 * n), o_s = A.add(sizeof(double) * 40 * n), o_c = A.add(sizeof(int32_t) * n);
  int rc = A.commit();
  if (rc != NWB_OK) return rc;
  cudaStream_t s = ctx->stream;
  S_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_t), targets, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, s));
  S_CUDA(ctx, cudaMemsetAsync(A.at<char>(o_s), 0xff, sizeof(double) * 40 * n, s));           // unused slots read back as NaN
  S_CUDA(ctx, launch_arm_ik(A.at<double>(o_t), n, A.at<double>(o_s), A.at<int32_t>(o_c), s));
  S_CUDA(ctx, cudaMemcpyAsync(solutions, A.at<double>(o_s), sizeof(double) * 40 * n, cudaMemcpyDeviceToHost, s));
  S_CUDA(ctx, cudaMemcpyAsync(counts, A.at<int32_t>(o_c), sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  S_CUDA(ctx, cudaStreamSynchronize(s));
  ctx->last_launches = 1;
  return NWB_OK;
}

int nwb_goal_arm_ik_batch(nwb_ctx* ctx, const double* q, int64_t n, const double* hand_pose6, uint64_t seed, uint64_t first_index,
                          uint8_t* ok, double* arm, double* ergonomics) {
  if (!ctx) return NWB_E_INVALID;
  if (n < 0 || !hand_pose6 || (n > 0 && (!q || !ok || !arm || !ergonomics))) return bad(ctx, "nwb_goal_arm_ik_batch: bad argument");
  if (n == 0) return NWB_OK;
  S_CUDA(ctx, cudaSetDevice(ctx->device));
  Scratch A(ctx);
  size_t o_q = A.add(sizeof(double) * NWB_NJ * n), o_ok = A.add(n), o_a = A.add(sizeof(double) * 5 * n), o_e = A.add(sizeof(double) * n);
  int rc = A.commit();
  if (rc != NWB_OK) return rc;
  cudaStream_t s = ctx->stream;
  PlanConst pc;
  make_plan_const(ctx, nullptr, ctx->params.planner_step_factor, pc);
  S_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_q), q, sizeof(double) * NWB_NJ * n, cudaMemcpyHostToDevice, s));
  S_CUDA(ctx, launch_goal_arm_ik(pc, hand_pose6, seed, first_index, A.at<double>(o_q), n, A.at<uint8_t>(o_ok), A.at<double>(o_a),
                                 A.at<double>(o_e), s));
  S_CUDA(ctx, cudaMemcpyAsync(ok, A.at<uint8_t>(o_ok), n, cudaMemcpyDeviceToHost, s));
  S_CUDA(ctx, cudaMemcpyAsync(arm, A.at<double>(o_a), sizeof(double) * 5 * n, cudaMemcpyDeviceToHost, s));
  S_CUDA(ctx, cudaMemcpyAsync(ergonomics, A.at<double>(o_e), sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  S_CUDA(ctx, cudaStreamSynchronize(s));
  ctx->last_launches = 1;
  return NWB_OK;
}

int nwb_hand_trajectory(int32_t kind, const double* start6, const double* goal6, int32_t num_interp_points, const double* door4,
                        double* out) {
  if (!start6 || !goal6 || !out || num_interp_points < 0 || (kind != 0 && kind != 1) || (kind == 1 && !door4)) return NWB_E_INVALID;
  if (kind == 0) drawer_points(start6, goal6, num_interp_points, out);
  else door_points(start6, goal6, num_interp_points, door4, out);
  return NWB_OK;
}

int nwb_enforce_ma_batch(nwb_ctx* ctx, const double* pts, int32_t n_points, int32_t start_tree, const int32_t* near_hand_index,
                         const double* q_near, const double* q_new_ds, int64_t n, double* q_out, int32_t* status) {
  if (!ctx) return NWB_E_INVALID;
  if (n < 0 || !pts || n_points < 2 || (n > 0 && (!near_hand_index || !q_near || !q_new_ds || !q_out || !status)))
    return bad(ctx, "nwb_enforce_ma_batch: bad argument");
  for (int64_t i = 0; i < n; ++i)
    if (near_hand_index[i] < 0 || near_hand_index[i] >= n_points) return bad(ctx, "nwb_enforce_ma_batch: hand index out of range");
  if (n == 0) return NWB_OK;
  S_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t qb = sizeof(double) * NWB_NJ * n, ib = sizeof(int32_t) * n, pb = sizeof(double) * 6 * n_points;
  Scratch A(ctx);
  size_t o_p = A.add(pb), o_h = A.add(ib), o_n = A.add(qb), o_d = A.add(qb), o_o = A.add(qb), o_s = A.add(ib);
  int rc = A.commit();
  if (rc != NWB_OK) return rc;
  cudaStream_t s = ctx->stream;
  PlanConst pc;
  make_plan_const(ctx, nullptr, ctx->params.planner_step_factor, pc);
  S_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_p), pts, pb, cudaMemcpyHostToDevice, s));
  S_CUDA(ctx, cudaMemcpyAsync(A.at<int32_t>(o_h), near_hand_index, ib, cudaMemcpyHostToDevice, s));
  S_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_n), q_near, qb, cudaMemcpyHostToDevice, s));
  S_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_d), q_new_ds, qb, cudaMemcpyHostToDevice, s));
  S_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_o), q_out, qb, cudaMemcpyHostToDevice, s));   // TRAPPED rows stay as the caller left them
  S_CUDA(ctx, launch_enforce_ma(pc, A.at<double>(o_p), n_points, start_tree, A.at<int32_t>(o_h), A.at<double>(o_n), A.at<double>(o_d), n,
                                A.at<double>(o_o), A.at<int32_t>(o_s), s));
  S_CUDA(ctx, cudaMemcpyAsync(q_out, A.at<double>(o_o), qb, cudaMemcpyDeviceToHost, s));
  S_CUDA(ctx, cudaMemcpyAsync(status, A.at<int32_t>(o_s), ib, cudaMemcpyDeviceToHost, s));
  S_CUDA(ctx, cudaStreamSynchronize(s));
  ctx->last_launches = 1;
  return NWB_OK;
}

int nwb_goal_sample_batch(nwb_ctx* ctx, uint64_t seed, uint64_t first, int64_t count, const double* hand_pose6, int32_t max_ik_iter,
                          uint8_t* stage, double* ergonomics, double* rows) {
  if (!ctx) return NWB_E_INVALID;
  if (count < 0 || !hand_pose6 || (count > 0 && !stage)) return bad(ctx, "nwb_goal_sample_batch: bad argument");
  if (count == 0) return NWB_OK;
  S_CUDA(ctx, cudaSetDevice(ctx->device));
  Scratch A(ctx);
  size_t o_rows = A.add(sizeof(double) * NWB_NJ * 16), o_idx = A.add(sizeof(uint64_t) * 16), o_erg = A.add(sizeof(double) * 16),
         o_n = A.add(8), o_st = A.add(count), o_ea = A.add(sizeof(double) * count), o_ra = A.add(sizeof(double) * NWB_NJ * count),
         o_pose = A.add(sizeof(double) * 6), o_seed = A.add(8);
  int rc = A.commit();
  if (rc != NWB_OK) return rc;
  cudaStream_t s = ctx->stream;
  PlanConst pc;
  make_plan_const(ctx, nullptr, ctx->params.planner_step_factor, pc);
  S_CUDA(ctx, cudaMemsetAsync(A.at<char>(o_n), 0, 8, s));
  S_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_pose), hand_pose6, sizeof(double) * 6, cudaMemcpyHostToDevice, s));
  S_CUDA(ctx, cudaMemcpyAsync(A.at<uint64_t>(o_seed), &seed, 8, cudaMemcpyHostToDevice, s));
  S_CUDA(ctx, ev_record(ctx, ctx->ev0, s));
  S_CUDA(ctx, launch_goal_sample(pc, ctx->tab_group, static_pairs(ctx), A.at<double>(o_pose), A.at<uint64_t>(o_seed), 1, first, count,
                                 max_ik_iter, A.at<double>(o_rows), A.at<uint64_t>(o_idx), A.at<double>(o_erg), 16,
                                 A.at<unsigned long long>(o_n), A.at<uint8_t>(o_st), ergonomics ? A.at<double>(o_ea) : nullptr,
                                 rows ? A.at<double>(o_ra) : nullptr, nullptr, s));
  S_CUDA(ctx, ev_record(ctx, ctx->ev1, s));
  S_CUDA(ctx, cudaMemcpyAsync(stage, A.at<uint8_t>(o_st), count, cudaMemcpyDeviceToHost, s));
  if (ergonomics) S_CUDA(ctx, cudaMemcpyAsync(ergonomics, A.at<double>(o_ea), sizeof(double) * count, cudaMemcpyDeviceToHost, s));
  if (rows) S_CUDA(ctx, cudaMemcpyAsync(rows, A.at<double>(o_ra), sizeof(double) * NWB_NJ * count, cudaMemcpyDeviceToHost, s));
  S_CUDA(ctx, cudaStreamSynchronize(s));
  ctx->last_ms = -1;
  ctx->last_launches = 1;
  return NWB_OK;
}

int nwb_sample_goal_configs(nwb_ctx* ctx, uint64_t seed, const double* hand_pose6, int32_t max_samples, int32_t max_ik_iter,
                            double* rows6, double* ergonomics6, uint64_t* index6, int32_t* n_found, int64_t* samples_drawn) {
  if (!ctx) return NWB_E_INVALID;
  if (!hand_pose6 || !rows6 || !n_found || max_samples < 0) return bad(ctx, "nwb_sample_goal_configs: bad argument");
  GoalSet g;
  int rc = sample_goals(ctx, seed, hand_pose6, max_samples, max_ik_iter, g);
  if (rc != NWB_OK) return rc;
  *n_found = g.n_found;
  if (samples_drawn) *samples_drawn = g.drawn;
  for (int r = 0; r < g.n_found; ++r) {
    std::memcpy(rows6 + (size_t)r * NWB_NJ, g.rows[r], sizeof(double) * NWB_NJ);
    if (ergonomics6) ergonomics6[r] = g.erg[r];
    if (index6) index6[r] = g.index[r];
  }
  return NWB_OK;
}

int nwb_time_parameterize(const nwb_ctx* ctx, const double* positions, int32_t n_points, double* time_from_start) {
  if (!ctx || n_points < 0 || (n_points > 0 && (!positions || !time_from_start))) return NWB_E_INVALID;
  iptp::time_stamps(ctx->params, positions, n_points, time_from_start);
  return NWB_OK;
}

int nwb_ctx_set_ds_database(nwb_ctx* ctx, const double* rows, int64_t n_rows) {
  if (!ctx) return NWB_E_INVALID;
  if (n_rows < 0 || (n_rows > 0 && !rows)) return bad(ctx, "nwb_ctx_set_ds_database: bad argument");
  ctx->ds_db.assign(rows, rows + (size_t)n_rows * NWB_NJ);
  ++ctx->ds_db_version;
  return NWB_OK;
}

int nwb_ctx_load_ds_database(nwb_ctx* ctx, const char* path) {
  if (!ctx) return NWB_E_INVALID;
  if (!path) return bad(ctx, "nwb_ctx_load_ds_database: bad argument");
  int64_t n = 0;
  int rc = nwb_load_ds_database(path, nullptr, 0, &n);
  if (rc != NWB_OK) {
    ctx->err = std::string("nwb_ctx_load_ds_database: cannot open ") + path;
    return rc;
  }
  std::vector<double> rows((size_t)std::max<int64_t>(n, 1) * NWB_NJ);
  rc = nwb_load_ds_database(path, rows.data(), n, &n);
  if (rc != NWB_OK) return rc;
  rows.resize((size_t)n * NWB_NJ);
  ctx->ds_db.swap(rows);
  ++ctx->ds_db_version;
  return NWB_OK;
}

int64_t nwb_ctx_ds_database_size(const nwb_ctx* ctx) { return ctx ? (int64_t)(ctx->ds_db.size() / NWB_NJ) : 0; }

int nwb_compute_goal_config(nwb_ctx* ctx, const nwb_compute_goal_config_request* req, nwb_compute_goal_config_response* resp,
                            uint64_t seed) {
  if (!ctx) return NWB_E_INVALID;
  if (!req || !resp) return bad(ctx, "nwb_compute_goal_config: bad argument");
  double pose6[6];
  pose_of(req->right_hand_position, req->right_hand_direction, pose6);
  GoalSet g;
  int rc = sample_goals(ctx, seed, pose6, ctx->params.scg_max_samples, ctx->params.scg_max_ik_iterations, g);
  if (rc != NWB_OK) return rc;
  resp->num_goal_configs_found = g.n_found;
  std::memset(resp->wb_goal_config, 0, sizeof resp->wb_goal_config);
  if (g.n_found) std::memcpy(resp->wb_goal_config, g.rows[0], sizeof(double) * NWB_NJ);
  return NWB_OK;
}

int nwb_compute_motion_plan(nwb_ctx* ctx, const nwb_compute_motion_plan_request* req, nwb_compute_motion_plan_response* resp,
                            uint64_t seed) {
  if (!ctx) return NWB_E_INVALID;
  if (!req || !resp || !req->wb_start_config || req->wb_start_config_len < NWB_NJ) return bad(ctx, "nwb_compute_motion_plan: bad argument");
  const nwb_params& P = ctx->params;
  double pose6[6];
  pose_of(req->right_hand_position, req->right_hand_direction, pose6);
  resp->motion_plan.n_points = 0;
  resp->num_nodes_generated = 0;
  resp->search_time = 0.0;
  resp->success = 0;
  GoalSet g;
  const double t0 = now_s();
  int rc = sample_goals(ctx, seed, pose6, P.scg_max_samples, P.scg_max_ik_iterations, g);
  if (rc != NWB_OK) return rc;
  resp->time_goal_config = now_s() - t0;
  if (g.n_found == 0) return NWB_OK;                                      // NPC:331-334
  std::memcpy(resp->wb_goal_config, g.rows[0], sizeof(double) * NWB_NJ);
  double w[NWB_NJ];
  limb_weights(P.r_arm_weight_approach, P.l_arm_weight_approach, P.legs_weight_approach, w);
  StageOut st;
  rc = plan_stage(ctx, req->wb_start_config, g.rows[0], w, seed + 2, nullptr, 0, req->execute_trajectory != 0, st);
  if (rc != NWB_OK) return rc;
  resp->num_nodes_generated = st.num_nodes;
  resp->search_time = st.search_time;
  resp->success = st.success;
  if (st.success) put_trajectory(P, st.path, false, req->execute_trajectory != 0, 0.f, resp->motion_plan);
  return NWB_OK;
}

int nwb_compute_motion_plan_batch(nwb_ctx* ctx, const nwb_compute_motion_plan_request* reqs, nwb_compute_motion_plan_response* resps,
                                  int64_t n, const uint64_t* seeds) {
  if (!ctx) return NWB_E_INVALID;
  if (n < 0 || (n > 0 && (!reqs || !resps || !seeds))) return bad(ctx, "nwb_compute_motion_plan_batch: bad argument");
  for (int64_t k = 0; k < n; ++k)
    if (!reqs[k].wb_start_config || reqs[k].wb_start_config_len < NWB_NJ) return bad(ctx, "nwb_compute_motion_plan_batch: bad argument");
  if (n == 0) return NWB_OK;
  if (ctx->ds_db.size() < 2 * NWB_NJ) {
    ctx->err = "no stable-configuration database loaded (nwb_ctx_set_ds_database / nwb_ctx_load_ds_database)";
    return NWB_E_STATE;
  }
  const nwb_params& P = ctx->params;
  const int N = (int)n;
  std::vector<double> poses((size_t)N * 6);
  for (int k = 0; k < N; ++k) pose_of(reqs[k].right_hand_position, reqs[k].right_hand_direction, &poses[(size_t)k * 6]);
  std::vector<GoalSet> goals(N);
  double t0 = now_s();
  int rc = sample_goals_batch(ctx, N, seeds, poses.data(), P.scg_max_samples, P.scg_max_ik_iterations, goals.data());
  if (rc != NWB_OK) return rc;
  const double t_goal = now_s() - t0;
  std::vector<int> live;                                    // requests with a goal configuration (NPC:270)
  for (int k = 0; k < N; ++k) {
    resps[k].motion_plan.n_points = 0;
    resps[k].num_nodes_generated = 0;
    resps[k].search_time = 0.0;
    resps[k].success = 0;
    resps[k].time_goal_config = t_goal;
    if (goals[k].n_found) {
      std::memcpy(resps[k].wb_goal_config, goals[k].rows[0], sizeof(double) * NWB_NJ);
      live.push_back(k);
    }
  }
  const int Q = (int)live.size();
  if (Q == 0) return NWB_OK;
  double w[NWB_NJ];
  limb_weights(P.r_arm_weight_approach, P.l_arm_weight_approach, P.legs_weight_approach, w);
  const int cap = 2048;
  std::vector<double> starts((size_t)Q * NWB_NJ), gcfg((size_t)Q * NWB_NJ);
  std::unique_ptr<double[]> raw(new double[(size_t)Q * cap * NWB_NJ]);    // not zero-filled: only the path rows are ever touched
  std::vector<uint64_t> qseeds(Q);
  std::vector<nwb_plan_result> res(Q);
  for (int i = 0; i < Q; ++i) {
    std::memcpy(&starts[(size_t)i * NWB_NJ], reqs[live[i]].wb_start_config, sizeof(double) * NWB_NJ);
    std::memcpy(&gcfg[(size_t)i * NWB_NJ], goals[live[i]].rows[0], sizeof(double) * NWB_NJ);
    qseeds[i] = seeds[live[i]] + 2;
  }
  t0 = now_s();
  rc = nwb_solve_query_batch(ctx, starts.data(), gcfg.data(), Q, ctx->ds_db.data(), (int32_t)(ctx->ds_db.size() / NWB_NJ), w,
                             qseeds.data(), P.planner_max_iterations, P.planner_step_factor, 0, res.data(), raw.get(), cap);
  if (rc == NWB_E_NOMEM) rc = NWB_OK;       // a search that outgrew the largest forest is that request's failure, not the batch's
  if (rc != NWB_OK) return rc;
  const double t_search = now_s() - t0;
  std::vector<std::vector<double>> paths;
  std::vector<int> owner;
  for (int i = 0; i < Q; ++i) {
    resps[live[i]].search_time = t_search;
    if (!res[i].success) continue;
    resps[live[i]].num_nodes_generated = (uint32_t)res[i].num_nodes_generated;
    const int len = res[i].raw_path_len;
    if (len > cap) {                        // longer than the shared buffer's slot: fetch this path whole, never truncate
      std::vector<double> whole((size_t)len * NWB_NJ);
      nwb_plan_result r1;
      rc = nwb_solve_query_batch(ctx, &starts[(size_t)i * NWB_NJ], &gcfg[(size_t)i * NWB_NJ], 1, ctx->ds_db.data(),
                                 (int32_t)(ctx->ds_db.size() / NWB_NJ), w, &qseeds[i], P.planner_max_iterations,
                                 P.planner_step_factor, 0, &r1, whole.data(), len);
      if (rc != NWB_OK) return rc;
      if (!r1.success || r1.raw_path_len != len) return bad(ctx, "nwb_compute_motion_plan_batch: a repeated search differs");
      paths.push_back(std::move(whole));
    } else {
      paths.emplace_back(raw.get() + (size_t)i * cap * NWB_NJ, raw.get() + ((size_t)i * cap + len) * NWB_NJ);
    }
    owner.push_back(live[i]);
  }
  std::vector<std::vector<int>> sel;
  std::vector<uint8_t> ok;
  t0 = now_s();
  rc = shortcut_paths_batch(ctx, paths, sel, ok);
  if (rc != NWB_OK) return rc;
  const double t_shortcut = now_s() - t0;
  t0 = now_s();
  struct Trace {                                            // NWB_TRACE=1: where a batch spends its time, on stderr
    double goal, search, shortcut, &start;
    int n, q;
    ~Trace() {
      if (trace_on())
        std::fprintf(stderr, "[nwb] compute_motion_plan_batch n=%d goals=%d: goal sampling %.3f ms, search %.3f ms, shortcut %.3f ms, trajectories %.3f ms\n",
                     n, q, goal * 1e3, search * 1e3, shortcut * 1e3, (now_s() - start) * 1e3);
    }
  } trace{t_goal, t_search, t_shortcut, t0, N, Q};
  parallel_for((int)paths.size(), [&](int i) {
    if (!ok[i]) return;                                     // shortcutting failed: empty trajectory (RP:1919-1923)
    const int k = owner[i];
    std::vector<double> way;
    for (int idx : sel[i]) way.insert(way.end(), paths[i].begin() + (size_t)idx * NWB_NJ, paths[i].begin() + (size_t)(idx + 1) * NWB_NJ);
    std::vector<double> path;
    const int n_sel = (int)sel[i].size();
    if (!reqs[k].execute_trajectory) {
      const int rows = (n_sel - 1) * (P.shortcut_waypoints + 1) + 1;
      path.resize((size_t)rows * NWB_NJ);
      nwb_interpolate_path(way.data(), n_sel, P.shortcut_waypoints, path.data(), rows);
    } else {
      path = way;
    }
    resps[k].success = !path.empty();
    put_trajectory(P, path, false, reqs[k].execute_trajectory != 0, 0.f, resps[k].motion_plan);
  });
  return NWB_OK;
}

int nwb_compute_linear_manipulation_plan(nwb_ctx* ctx, const nwb_compute_linear_manipulation_plan_request* req,
                                         nwb_manipulation_plan_response* resp, uint64_t seed) {
  if (!ctx) return NWB_E_INVALID;
  if (!req || !resp || !req->wb_start_config || req->wb_start_config_len < NWB_NJ)
    return bad(ctx, "nwb_compute_linear_manipulation_plan: bad argument");
  double pose6[6];
  pose_of(req->right_hand_position, req->right_hand_direction, pose6);
  const double object[2] = {req->relative_angle, req->object_translation_length};
  return manipulation_plan(ctx, pose6, req->wb_start_config, req->execute_trajectory != 0, false, object, resp, seed);
}

int nwb_compute_circular_manipulation_plan(nwb_ctx* ctx, const nwb_compute_circular_manipulation_plan_request* req,
                                           nwb_manipulation_plan_response* resp, uint64_t seed) {
  if (!ctx) return NWB_E_INVALID;
  if (!req || !resp || !req->wb_start_config || req->wb_start_config_len < NWB_NJ)
    return bad(ctx, "nwb_compute_circular_manipulation_plan: bad argument");
  double pose6[6];
  pose_of(req->right_hand_position, req->right_hand_direction, pose6);
  const double object[4] = {req->relative_angle, req->rotation_radius, req->rotation_angle, req->clearance_handle};
  return manipulation_plan(ctx, pose6, req->wb_start_config, req->execute_trajectory != 0, true, object, resp, seed);
}

}  // extern "C"
