// TEST INFRASTRUCTURE - CPU oracle, never shipped, never on the product path.
//
// Function-for-function fp64 restatement of the reference planner's hot path.  File:line
// citations are relative to the reference checkout:
//   RP  = rrt_connect_planner/src/rrt_planner.cpp
//   LP  = rrt_connect_planner/src/local_planner.cpp
//   SCG = rrt_connect_planner/src/stable_config_generator.cpp
//   NCK = rrt_connect_planner/src/nao_constraint_kinematics.cpp
// The reference seeds its RNG from the clock (random_numbers::RandomNumberGenerator, RPh:142), so
// "identical results" is only defined with an injected stream: a counter-based Philox4x32-10
// shared (as a specification, not as code) with the product - see DESIGN.md "random streams".
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "../include/nwb.h"
#include "model_eval.h"

namespace oracle {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11).
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1,
             n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
constexpr uint32_t kStreamSampler = 1, kStreamPlanner = 2;
inline double u01(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }
// uniform number `slot` of item `index` in stream `stream`
inline double stream_uniform(uint64_t seed, uint32_t stream, uint64_t index, uint32_t slot) {
  uint32_t ctr[4] = {(uint32_t)index, (uint32_t)(index >> 32), slot / 4, stream};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  philox4x32_10(ctr, key, out);
  return u01(out[slot % 4]);
}

// ---------------------------------------------------------------------------------------------
// The hand-built KDL chains (LP:18-47, NCK:23-47).
struct Chains {
  kdlr::Chain right_leg, left_leg, legs, right_arm;
  Chains() {
    using namespace kdlr;
    auto rl = [](Chain& c) {
      c.add(None, Vec{0, 0, 0.04519});
      c.add(RotX, Vec{0, 0, 0});
      c.add(RotY, Vec{0, 0, 0.1029});
      c.add(RotY, Vec{0, 0, 0.1});
      c.add(RotY, Vec{0, 0, 0});
      c.add(RotX, Vec{0, 0.05, 0});
    };
    auto ll = [](Chain& c) {
      c.add(None, Vec{0, 0.05, 0});
      c.add(RotX, Vec{0, 0, 0});
      c.add(RotY, Vec{0, 0, -0.1});
      c.add(RotY, Vec{0, 0, -0.1029});
      c.add(RotY, Vec{0, 0, 0});
      c.add(RotX, Vec{0, 0, -0.04519});
    };
    rl(right_leg);
    ll(left_leg);
    rl(legs);
    ll(legs);
    // NCK:39-47
    right_arm.add(None, Vec{0, 0, 0.182});
    right_arm.add(None, Vec{0, -0.098, 0});
    right_arm.add(RotY, Vec{0, 0, 0});
    right_arm.add(RotZ, Vec{0.105, -0.015, 0});
    right_arm.add(RotX, Vec{0, 0, 0});
    right_arm.add(RotZ, Vec{0.05595, 0, 0});
    right_arm.add(RotX, Vec{0.05775, 0, -0.01231});
    right_arm.add(None, Frame{rot_rpy(M_PI_2, 0, 0), Vec{0.01, 0, 0}});
  }
};
inline const Chains& chains() {
  static const Chains c;
  return c;
}

struct PlanParams {
  EvalParams eval;
  unsigned quirks = NWB_QUIRKS_DEFAULT;
  double ds_tolerance = 0.01;
  double ik_eps = 1e-3;
  int ik_maxiter = 100;
  double pinv_eps = 1e-5;
};

// LocalPlanner::checkJointLimits (LP:305-371) / ConstraintKinematics::joint_limits_respected
// (NCK:549-608): strict inequalities, swapped bounds tolerated.
inline bool check_joint_limits(const double* cfg, const double* mn, const double* mx, int n) {
  bool valid = true;
  for (int i = 0; i < n; ++i) {
    if (mn[i] < mx[i]) {
      if (cfg[i] > mn[i]) {
        if (!(cfg[i] < mx[i])) valid = false;
      } else
        valid = false;
    }
    if (mx[i] < mn[i]) {
      if (cfg[i] > mx[i]) {
        if (!(cfg[i] < mn[i])) valid = false;
      } else
        valid = false;
    }
  }
  return valid;
}

// LocalPlanner::getDesiredLeftLegPose (LP:59-110) == first half of NCK:71-115:
// target = (R^T, R^T ((0,0.1,0) - p_hip)).
inline Frame desired_left_leg_pose(const Frame& hip) {
  Frame t;
  t.M = kdlr::transpose(hip.M);
  Vec fixed{nwbm::kLeftSoleTarget[0], nwbm::kLeftSoleTarget[1], nwbm::kLeftSoleTarget[2]};
  for (int i = 0; i < 3; ++i) {
    double tmp = -t.M(i, 0) * hip.p.x - t.M(i, 1) * hip.p.y - t.M(i, 2) * hip.p.z;
    double tmp2 = t.M(i, 0) * fixed.x + t.M(i, 1) * fixed.y + t.M(i, 2) * fixed.z;
    (i == 0 ? t.p.x : i == 1 ? t.p.y : t.p.z) = tmp + tmp2;
  }
  return t;
}

// LocalPlanner::checkLeftLegPose (LP:244-299).  Quirk C2: the shipped code takes the signed error
// (desired - actual) and its max starting at 0; orientation is never looked at.
inline bool check_left_leg_pose(const PlanParams& pp, const double* legs_config /*[10]*/) {
  Frame lf = kdlr::fk(chains().legs, legs_config);
  double des[3] = {nwbm::kLeftSoleTarget[0], nwbm::kLeftSoleTarget[1], nwbm::kLeftSoleTarget[2]};
  double act[3] = {lf.p.x, lf.p.y, lf.p.z};
  double max_err = 0.0;
  for (int i = 0; i < 3; ++i) {
    double e = des[i] - act[i];
    if (!(pp.quirks & NWB_QUIRK_SIGNED_FOOT_ERROR)) e = std::fabs(e);
    if (e > max_err) max_err = e;
  }
  return !(max_err > pp.ds_tolerance);
}

// LocalPlanner::computeLeftLegIK (LP:151-236): seed = current left leg.
inline bool left_leg_ik_local(const PlanParams& pp, const double* legs_config /*[10]*/, const double* mn5,
                              const double* mx5, double* out5, int* iters = nullptr) {
  Frame hip = kdlr::fk(chains().right_leg, legs_config);
  Frame target = desired_left_leg_pose(hip);
  double q_out[5];
  int rc = kdlr::ik_pos_nr(chains().left_leg, legs_config + 5, target, q_out, pp.ik_maxiter, pp.ik_eps, pp.pinv_eps, iters);
  if (rc < 0) return false;
  if (!check_joint_limits(q_out, mn5, mx5, 5)) return false;
  for (int i = 0; i < 5; ++i) out5[i] = q_out[i];
  return true;
}

// ConstraintKinematics::getHipPose + computeLeftLegIK (NCK:512-544, NCK:71-197): seed
// q_init(i) = -r_leg_config_[i] (NCK:150) where r_leg_config_ = -q[0..4]  =>  seed = q[0..4]
// applied index-wise; every one of the max_ik_iter attempts is identical (quirk C5).
inline bool left_leg_ik_sampler(const PlanParams& pp, const double* right_leg_values /*[5] = -q[0..4]*/,
                                const double* mn5, const double* mx5, int max_ik_iter, double* out5,
                                int* iters = nullptr) {
  Frame hip = kdlr::fk(chains().right_leg, right_leg_values);
  Frame target = desired_left_leg_pose(hip);
  double q_init[5], q_out[5];
  int rc = -3;
  bool valid = false;
  for (int it = 0; it < max_ik_iter; ++it) {
    for (int i = 0; i < 5; ++i) q_init[i] = -right_leg_values[i];
    rc = kdlr::ik_pos_nr(chains().left_leg, q_init, target, q_out, pp.ik_maxiter, pp.ik_eps, pp.pinv_eps, iters);
    if (rc >= 0) {
      valid = check_joint_limits(q_out, mn5, mx5, 5);
      if (valid) break;
    }
  }
  if (rc >= 0 && valid) {
    for (int i = 0; i < 5; ++i) out5[i] = q_out[i];
    return true;
  }
  return false;
}

// RRT_Planner::generate_q_new (RP:424-494).
inline int generate_q_new(const double* q_near, const double* q_rand, const double* w, double step, double* q_new) {
  const int n = NWB_NJ;
  double d[NWB_NJ], sum = 0;
  for (int i = 0; i < n; ++i) {
    d[i] = q_rand[i] - q_near[i];
    sum = sum + d[i] * d[i];
  }
  double norm = std::sqrt(sum);
  double inc[NWB_NJ], sum_inc = 0;
  for (int j = 0; j < n; ++j) {
    double dir = d[j] / norm;
    inc[j] = step * (dir * w[j]);
    sum_inc = sum_inc + inc[j] * inc[j];
  }
  double norm_inc = std::sqrt(sum_inc);
  if (norm_inc < norm) {
    for (int i = 0; i < n; ++i) q_new[i] = q_near[i] + inc[i];
    return NWB_ADVANCED;
  }
  for (int i = 0; i < n; ++i) q_new[i] = q_rand[i];
  return NWB_REACHED;
}

// RRT_Planner::enforce_DS_Constraints (RP:498-593).
inline int enforce_ds_constraints(const PlanParams& pp, const double* q_new, double* q_mod) {
  if (!check_joint_limits(q_new, nwbm::kJointMin, nwbm::kJointMax, NWB_NJ)) return NWB_TRAPPED;
  double legs[10];
  for (int i = 0; i < 5; ++i) legs[i] = -q_new[i];
  for (int i = 16; i < NWB_NJ; ++i) legs[i - 11] = q_new[i];
  if (check_left_leg_pose(pp, legs)) {
    for (int i = 0; i < NWB_NJ; ++i) q_mod[i] = q_new[i];
    return NWB_REACHED;
  }
  double sol[5];
  if (!left_leg_ik_local(pp, legs, nwbm::kJointMin + 16, nwbm::kJointMax + 16, sol)) return NWB_TRAPPED;
  for (int i = 0; i < NWB_NJ; ++i) q_mod[i] = q_new[i];
  for (int i = 16; i < NWB_NJ; ++i) q_mod[i] = sol[i - 16];
  return NWB_ADVANCED;
}

// RRT_Planner::isConfigValid (RP:701-745).
inline bool is_config_valid(const PlanParams& pp, const double* q) { return is_feasible(pp.eval, q, nullptr, nullptr); }

struct Node {  // RPh:30-36 (cart_hand_pose_index unused without an articulation constraint)
  int index, predecessor_index;
  double config[NWB_NJ];
};
struct Tree {
  std::vector<Node> nodes;
};

// RRT_Planner::findNearestNeighbour (RP:336-419): strict <, initial best 10000.0 -> first minimum
// wins.  Returns -1 when no node is closer than 1e4 (quirk C4).
inline int nearest_neighbour(const double* nodes /*[N][21]*/, int64_t N, const double* q, double* best_out) {
  int64_t nn = -1;
  double best = 10000.0;
  for (int64_t i = 0; i < N; ++i) {
    double sum = 0;
    for (int j = 0; j < NWB_NJ; ++j) {
      double d = q[j] - nodes[i * NWB_NJ + j];
      sum = sum + d * d;
    }
    double dn = std::sqrt(sum);
    if (dn < best) {
      best = dn;
      nn = i;
    }
  }
  if (best_out) *best_out = best;
  return (int)nn;
}
inline int nearest_in_tree(const Tree& t, const double* q) {
  int nn = -1;
  double best = 10000.0;
  for (size_t i = 0; i < t.nodes.size(); ++i) {
    double sum = 0;
    for (int j = 0; j < NWB_NJ; ++j) {
      double d = q[j] - t.nodes[i].config[j];
      sum = sum + d * d;
    }
    double dn = std::sqrt(sum);
    if (dn < best) {
      best = dn;
      nn = (int)i;
    }
  }
  return nn;
}

inline void add_config(Tree& t, int near_index, const double* q) {  // RP:749-783
  Node n;
  n.index = (int)t.nodes.size();
  n.predecessor_index = near_index;
  for (int i = 0; i < NWB_NJ; ++i) n.config[i] = q[i];
  t.nodes.push_back(n);
}

// RRT_Planner::extendTree (RP:787-884), articulation constraint inactive.
inline int extend_tree(const PlanParams& pp, Tree& t, const double* q_rand, const double* w, double step, int* q_near_idx) {
  int nn = nearest_in_tree(t, q_rand);
  *q_near_idx = nn;
  double q_new[NWB_NJ], q_mod[NWB_NJ];
  int state = generate_q_new(t.nodes[nn].config, q_rand, w, step, q_new);
  int ds = enforce_ds_constraints(pp, q_new, q_mod);
  if (ds == NWB_TRAPPED) return NWB_TRAPPED;
  if (ds == NWB_ADVANCED) state = NWB_ADVANCED;
  if (!is_config_valid(pp, q_mod)) return NWB_TRAPPED;
  add_config(t, t.nodes[nn].index, q_mod);
  return state;
}

// RRT_Planner::connectTree (RP:890-1034), articulation constraint inactive.  *q_near_idx is the
// node that finally connected (RP:1000-1001) or the initial nearest neighbour.
inline int connect_tree(const PlanParams& pp, Tree& t, const double* q_connect, const double* w, double step,
                        int* q_near_idx, int* steps_out = nullptr) {
  int nn = nearest_in_tree(t, q_connect);
  *q_near_idx = nn;
  int cur = nn;
  int state = NWB_ADVANCED, steps = 0;
  while (state != NWB_TRAPPED && state != NWB_REACHED) {
    double q_new[NWB_NJ], q_mod[NWB_NJ];
    state = generate_q_new(t.nodes[cur].config, q_connect, w, step, q_new);
    ++steps;
    int ds = enforce_ds_constraints(pp, q_new, q_mod);
    if (ds == NWB_TRAPPED) {
      state = NWB_TRAPPED;
      break;
    }
    if (ds == NWB_ADVANCED) state = NWB_ADVANCED;
    if (!is_config_valid(pp, q_mod)) {
      state = NWB_TRAPPED;
      break;
    }
    add_config(t, t.nodes[cur].index, q_mod);
    if (state == NWB_REACHED)
      *q_near_idx = cur;
    else
      cur = (int)t.nodes.size() - 1;
  }
  if (steps_out) *steps_out = steps;
  return state;
}

// RRT_Planner::interpolateWaypoints, arithmetic only (RP:1752-1774): k+2 states.
inline void interpolate_waypoints(const double* a, const double* b, int k, std::vector<double>& out) {
  out.resize((size_t)(k + 2) * NWB_NJ);
  double sw[NWB_NJ];
  for (int j = 0; j < NWB_NJ; ++j) sw[j] = (b[j] - a[j]) / (k + 1);
  for (int j = 0; j < NWB_NJ; ++j) out[j] = a[j];
  for (int n = 1; n <= k + 1; ++n)
    for (int j = 0; j < NWB_NJ; ++j) out[(size_t)n * NWB_NJ + j] = a[j] + sw[j] * n;
}

// PlanningScene::isPathValid as called at RP:1876,1905: every way-point must be collision-free
// over the whole robot; no feasibility predicate is registered (RP:233 commented) - quirk C6.
inline bool is_path_valid(const PlanParams& pp, const double* a, const double* b, int k, int* first_bad = nullptr) {
  std::vector<double> w;
  interpolate_waypoints(a, b, k, w);
  for (int n = 0; n < k + 2; ++n) {
    Frame fr[nwbm::kNumFrames];
    fk_all(&w[(size_t)n * NWB_NJ], fr);
    if (in_collision(pp.eval, fr, /*group_only=*/false, nullptr)) {
      if (first_bad) *first_bad = n;
      return false;
    }
  }
  if (first_bad) *first_bad = -1;
  return true;
}

// RRT_Planner::interpolateShortPathWaypoints (RP:1956-1994).
inline void interpolate_short_path(const std::vector<double>& path, int k, std::vector<double>& out) {
  const size_t W = path.size() / NWB_NJ;
  out.assign(path.begin(), path.begin() + NWB_NJ);
  for (size_t i = 0; i + 1 < W; ++i) {
    double sw[NWB_NJ];
    for (int j = 0; j < NWB_NJ; ++j) sw[j] = (path[(i + 1) * NWB_NJ + j] - path[i * NWB_NJ + j]) / (k + 1);
    for (int n = 1; n <= k + 1; ++n)
      for (int j = 0; j < NWB_NJ; ++j) out.push_back(path[i * NWB_NJ + j] + sw[j] * n);
  }
}

// RRT_Planner::path_shortcutter (RP:1830-1951), way-point selection only.  Returns false when the
// 500-round budget is exhausted.  *edges_checked counts isPathValid calls.
inline bool path_shortcutter(const PlanParams& pp, const std::vector<double>& raw, int k, int max_rounds,
                             std::vector<double>& out, int* edges_checked = nullptr) {
  const int W = (int)(raw.size() / NWB_NJ);
  auto row = [&](int i) { return &raw[(size_t)i * NWB_NJ]; };
  auto push = [&](const double* r) { out.insert(out.end(), r, r + NWB_NJ); };
  out.clear();
  push(row(0));
  std::vector<double> cur(row(0), row(0) + NWB_NJ);
  int cur_idx = 0, edges = 0, budget = max_rounds;
  bool reached = false;
  while (!reached && budget > 0) {
    ++edges;
    if (is_path_valid(pp, cur.data(), row(W - 1), k)) {
      push(row(W - 1));
      reached = true;
    } else {
      for (int i = W - 2; i >= 0; --i) {
        if (cur_idx == i) {
          cur.assign(row(i + 1), row(i + 1) + NWB_NJ);
          cur_idx = i + 1;
          push(cur.data());
          break;
        }
        ++edges;
        if (is_path_valid(pp, cur.data(), row(i), k)) {
          cur.assign(row(i), row(i) + NWB_NJ);
          cur_idx = i;
          push(row(i));
          break;
        }
      }
    }
    --budget;
  }
  if (edges_checked) *edges_checked = edges;
  return budget != 0;
}

struct SolveResult {
  bool success = false;
  int iterations = 0, num_nodes = 0, nodes_start = 0, nodes_goal = 0;
  std::vector<double> raw_path, short_path;
};

// RRT_Planner::getRandomStableConfig (RP:316-332) with the injected stream.
inline int random_db_index(unsigned quirks, uint64_t seed, uint64_t iteration, int n_db) {
  double u = stream_uniform(seed, kStreamPlanner, iteration, 0);
  if (quirks & NWB_QUIRK_DB_INDEX_FLOOR) return (int)std::floor(0.0 + (double)(n_db - 1 - 0) * u);
  return (int)std::floor((double)n_db * u);
}

// RRT_Planner::setStartGoalConfigs + solveQuery + writePath path extraction
// (RP:227-302, RP:1039-1212, RP:1303-1392), articulation constraint inactive.
inline SolveResult solve_query(const PlanParams& pp, const double* start, const double* goal, const double* db,
                               int n_db, const double* w, uint64_t seed, int max_iter, double step) {
  SolveResult res;
  if (!is_config_valid(pp, start) || !is_config_valid(pp, goal)) return res;
  Tree Ts, Tg;
  add_config(Ts, 0, start);
  add_config(Tg, 0, goal);
  bool swap = false;
  for (int it = 0; it < max_iter; ++it) {
    res.iterations = it + 1;
    const double* q_rand = db + (size_t)random_db_index(pp.quirks, seed, (uint64_t)it, n_db) * NWB_NJ;
    Tree& Ta = swap ? Tg : Ts;
    Tree& Tb = swap ? Ts : Tg;
    int q_near = -1;
    int returned = extend_tree(pp, Ta, q_rand, w, step, &q_near);
    if (returned != NWB_TRAPPED) {
      int found = connect_tree(pp, Tb, Ta.nodes.back().config, w, step, &q_near);
      if (found == NWB_REACHED) {
        // writePath: connector 1 = goal tree connected to start tree (RP:1317-1327)
        int idx_s = swap ? q_near : Ts.nodes.back().index;
        int idx_g = swap ? Tg.nodes.back().index : q_near;
        std::vector<const double*> sp, gp;
        while (idx_s > 0) {
          sp.push_back(Ts.nodes[idx_s].config);
          idx_s = Ts.nodes[idx_s].predecessor_index;
        }
        sp.push_back(Ts.nodes[idx_s].config);
        while (idx_g > 0) {
          gp.push_back(Tg.nodes[idx_g].config);
          idx_g = Tg.nodes[idx_g].predecessor_index;
        }
        gp.push_back(Tg.nodes[idx_g].config);
        for (int i = (int)sp.size() - 1; i >= 0; --i) res.raw_path.insert(res.raw_path.end(), sp[i], sp[i] + NWB_NJ);
        for (size_t i = 0; i < gp.size(); ++i) res.raw_path.insert(res.raw_path.end(), gp[i], gp[i] + NWB_NJ);
        res.success = true;
        res.nodes_start = (int)Ts.nodes.size();
        res.nodes_goal = (int)Tg.nodes.size();
        res.num_nodes = res.nodes_start + res.nodes_goal;
        return res;
      }
    }
    swap = !swap;  // trees are swapped every iteration regardless of outcome (RP:1125,1131,1186,1192)
  }
  return res;
}

// StableConfigGenerator::sample_stable_configs, database branch (SCG:278-480) for the sample
// indices [first, first+count).  Sample i draws its 21 joints from stream (seed, sampler, i), so
// the accepted set does not depend on how the index range is partitioned.
// stage_out[i - first] (optional): 0 = IK failed, 1 = infeasible, 2 = feasible (before quirk C1).
inline void sample_stable_configs(const PlanParams& pp, uint64_t seed, uint64_t first, uint64_t count, int max_ik_iter,
                                  std::vector<double>& rows, std::vector<uint64_t>& row_index,
                                  std::vector<uint8_t>* stage_out = nullptr) {
  for (uint64_t i = first; i < first + count; ++i) {
    double q[NWB_NJ];
    for (int j = 0; j < NWB_NJ; ++j) {
      double u = stream_uniform(seed, kStreamSampler, i, (uint32_t)j);
      q[j] = (nwbm::kJointMax[j] - nwbm::kJointMin[j]) * u + nwbm::kJointMin[j];
    }
    q[15] = 0.0;  // SCG:291 LHipYawPitch fixed
    double rleg[5], lleg[5];
    for (int k = 0; k < 5; ++k) rleg[k] = -q[k];
    uint8_t stage = 0;
    if (left_leg_ik_sampler(pp, rleg, nwbm::kJointMin + 16, nwbm::kJointMax + 16, max_ik_iter, lleg)) {
      for (int k = 0; k < 5; ++k) q[16 + k] = lleg[k];
      bool feasible = is_feasible(pp.eval, q, nullptr, nullptr);
      stage = feasible ? 2 : 1;
      bool keep = (pp.quirks & NWB_QUIRK_DB_KEEPS_INFEASIBLE) ? !feasible : feasible;  // SCG:452-456
      if (keep) {
        rows.insert(rows.end(), q, q + NWB_NJ);
        row_index.push_back(i);
      }
    }
    if (stage_out) stage_out->push_back(stage);
  }
}

}  // namespace oracle
