// TEST INFRASTRUCTURE - CPU oracle, never shipped, never on the product path.
//
// Function-for-function fp64 restatement of the reference planner's hot path.  File:line
// citations are relative to the reference checkout:
//   RP  = rrt_connect_planner/src/rrt_planner.cpp
//   LP  = rrt_connect_planner/src/local_planner.cpp
//   SCG = rrt_connect_planner/src/stable_config_generator.cpp
//   NCK = rrt_connect_planner/src/nao_constraint_kinematics.cpp
//   AOC = rrt_connect_planner/src/articulated_object_constraints.cpp
//   NPC = rrt_connect_planner/src/nao_planner_control.cpp
// The reference seeds its RNG from the clock (random_numbers::RandomNumberGenerator, RPh:142), so
// "identical results" is only defined with an injected stream: a counter-based Philox4x32-10
// shared (as a specification, not as code) with the product - see DESIGN.md "random streams".
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/nwb.h"
#include "arm_restated.h"
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
constexpr uint32_t kStreamSampler = 1, kStreamPlanner = 2, kStreamArmSeed = 3;
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
  kdlr::Chain right_leg, left_leg, legs, right_arm, r_leg_arm;
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
    // AOC:42-56: right sole -> hip -> right hand (grasp frame)
    rl(r_leg_arm);
    for (size_t i = 0; i < right_arm.segs.size(); ++i) r_leg_arm.segs.push_back(right_arm.segs[i]);
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

// ---------------------------------------------------------------------------------------------
// Articulated-object (manipulation) constraint: the hand must follow a discretised trajectory.
struct MaConstraint {
  int num_points = 0;             // num_interp_points + 2 (AOC:192-195)
  std::vector<double> pts;        // [num_points][6]: position, direction of the hand z-axis
  const double* pt(int i) const { return &pts[(size_t)i * 6]; }
};
// ArticulatedObjectConstraints::add_drawer_constraint + computeDrawerTrajectoryPoints (AOC:68-122).
inline MaConstraint drawer_constraint(const double* start6, const double* goal6, int num_interp) {
  MaConstraint ma;
  ma.num_points = num_interp + 2;
  ma.pts.resize((size_t)ma.num_points * 6);
  double step[6];
  for (int i = 0; i < 6; ++i) step[i] = (goal6[i] - start6[i]) / (num_interp + 1);
  for (int s = 0; s < ma.num_points; ++s)
    for (int i = 0; i < 6; ++i) ma.pts[(size_t)s * 6 + i] = start6[i] + (step[i] * s);
  return ma;
}
// add_door_constraint + computeDoorTrajectoryPoints (AOC:86-189); door = {radius, opening angle [deg],
// relative angle [deg], handle clearance}.
inline MaConstraint door_constraint(const double* start6, const double* goal6, int num_interp, const double* door) {
  MaConstraint ma;
  ma.num_points = num_interp + 2;
  ma.pts.resize((size_t)ma.num_points * 6);
  double relative_angle = door[2] * (M_PI / 180);
  double step_angle = (door[1] * (M_PI / 180)) / (num_interp + 1);
  double sd[3];
  for (int i = 0; i < 3; ++i) sd[i] = (goal6[3 + i] - start6[3 + i]) / (num_interp + 1);
  double x_door = start6[0] + door[3] * std::cos(relative_angle);
  double y_door = start6[1] - door[3] * std::sin(relative_angle);
  double z_door = start6[2];
  double x_hinge = x_door - door[0] * std::sin(relative_angle);
  double y_hinge = y_door - door[0] * std::cos(relative_angle);
  double z_hinge = z_door;
  for (int s = 0; s < ma.num_points; ++s) {
    double tmp_angle = ((step_angle * s) - relative_angle);
    x_door = x_hinge - door[0] * std::sin(tmp_angle);
    y_door = y_hinge + door[0] * std::cos(tmp_angle);
    z_door = z_hinge;
    tmp_angle = 1.5708 - tmp_angle;
    double* o = &ma.pts[(size_t)s * 6];
    o[0] = x_door - door[3] * std::sin(tmp_angle);
    o[1] = y_door - door[3] * std::cos(tmp_angle);
    o[2] = z_door;
    for (int i = 0; i < 3; ++i) o[3 + i] = start6[3 + i] + (sd[i] * s);
  }
  return ma;
}

// Hand target of a pose given w.r.t. the right sole, expressed in the solver's torso frame
// (NCK:226-282 == AOC:316-357): R = hip^T, pos = R (p - p_hip) - (0, 0, 0.085), dir = R d.
inline void hand_target_in_torso(const Frame& hip, const double* pose6, double* pos, double* dir, kdlr::Rot* rot_out = nullptr) {
  kdlr::Rot R = kdlr::transpose(hip.M);
  for (int i = 0; i < 3; ++i) {
    double tmp = -R(i, 0) * hip.p.x - R(i, 1) * hip.p.y - R(i, 2) * hip.p.z;
    double tmp2 = R(i, 0) * pose6[0] + R(i, 1) * pose6[1] + R(i, 2) * pose6[2];
    pos[i] = tmp + tmp2;
  }
  pos[2] = pos[2] - 0.085;
  for (int i = 0; i < 3; ++i) dir[i] = R(i, 0) * pose6[3] + R(i, 1) * pose6[4] + R(i, 2) * pose6[5];
  if (rot_out) *rot_out = R;
}

// ArticulatedObjectConstraints::right_hand_at_pose (AOC:204-262): FK of the sole->hand chain;
// position within 5 mm and z-axis within 0.007 per component.
inline bool right_hand_at_pose(const MaConstraint& ma, int pose_index, const double* r_leg5, const double* r_arm5) {
  double q[10];
  for (int i = 0; i < 5; ++i) q[i] = r_leg5[i];
  for (int i = 0; i < 5; ++i) q[5 + i] = r_arm5[i];
  Frame hand = kdlr::fk(chains().r_leg_arm, q);
  const double* t = ma.pt(pose_index);
  double e[6] = {std::fabs(hand.p.x - t[0]), std::fabs(hand.p.y - t[1]), std::fabs(hand.p.z - t[2]),
                 std::fabs(hand.M(0, 2) - t[3]), std::fabs(hand.M(1, 2) - t[4]), std::fabs(hand.M(2, 2) - t[5])};
  double max_pos = 0.0, max_dir = 0.0;
  for (int i = 0; i < 3; ++i)
    if (e[i] > max_pos) max_pos = e[i];
  for (int i = 3; i < 6; ++i)
    if (e[i] > max_dir) max_dir = e[i];
  if (max_pos > 0.005) return false;
  if (max_dir > 0.007) return false;
  return true;
}

// ArticulatedObjectConstraints::compute_right_arm_IK (AOC:313-449): of the solutions inside the
// (strict) joint limits take the one with the smallest L1 distance to the reference arm posture.
inline bool ma_right_arm_ik(const MaConstraint& ma, int pose_index, const double* r_leg5, const double* arm_ref5,
                            const double* mn5, const double* mx5, double* out5) {
  Frame hip = kdlr::fk(chains().right_leg, r_leg5);
  double pos[3], dir[3], sols[8 * 5];
  hand_target_in_torso(hip, ma.pt(pose_index), pos, dir);
  int n = arm5d_ik(pos, dir, sols);
  if (n == 0) return false;
  double best = 10000.0;
  int best_i = 1000;
  for (int s = 0; s < n; ++s) {
    if (!check_joint_limits(sols + s * 5, mn5, mx5, 5)) continue;
    double dist = 0.0;
    for (int i = 0; i < 5; ++i) dist = dist + std::fabs(arm_ref5[i] - sols[s * 5 + i]);
    if (dist < best) {
      best = dist;
      best_i = s;
    }
  }
  if (best < 1000) {
    for (int i = 0; i < 5; ++i) out5[i] = sols[best_i * 5 + i];
    return true;
  }
  return false;
}

// RRT_Planner::enforce_MA_Constraints (RP:597-696).  near_hand_index = q_near.cart_hand_pose_index.
inline int enforce_ma_constraints(const MaConstraint& ma, bool start_tree, int near_hand_index, const double* q_near,
                                  const double* q_new_ds, double* q_new_manip) {
  int desired;
  if (start_tree && near_hand_index <= ma.num_points - 2)
    desired = near_hand_index + 1;
  else if (!start_tree && near_hand_index > 0)
    desired = near_hand_index - 1;
  else
    desired = near_hand_index;
  double r_leg[5], r_arm[5];
  for (int i = 0; i < 5; ++i) r_leg[i] = -q_new_ds[i];
  for (int i = 10; i < 15; ++i) r_arm[i - 10] = q_new_ds[i];
  if (right_hand_at_pose(ma, desired, r_leg, r_arm)) {
    for (int i = 0; i < NWB_NJ; ++i) q_new_manip[i] = q_new_ds[i];
    return NWB_REACHED;
  }
  double sol[5];
  if (ma_right_arm_ik(ma, desired, r_leg, q_near + 10, nwbm::kJointMin + 10, nwbm::kJointMax + 10, sol)) {
    for (int i = 0; i < NWB_NJ; ++i) q_new_manip[i] = q_new_ds[i];
    for (int i = 10; i < 15; ++i) q_new_manip[i] = sol[i - 10];
    return NWB_ADVANCED;
  }
  return NWB_TRAPPED;
}

struct Node {  // RPh:30-36
  int index, predecessor_index;
  double config[NWB_NJ];
  int cart_hand_pose_index = 0;
};
struct Tree {
  std::vector<Node> nodes;
  bool start_tree = true;          // Tree::name == "start_tree" (RP:95-96)
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

// RRT_Planner::addConfigtoTree (RP:749-783).  Note the bound `< num_points - 2` here against
// `<= num_points - 2` in enforce_MA_Constraints (RP:607): kept as shipped.
inline void add_config(Tree& t, int near_index, const double* q, const MaConstraint* ma = nullptr, int near_hand_index = 0) {
  Node n;
  n.index = (int)t.nodes.size();
  n.predecessor_index = near_index;
  for (int i = 0; i < NWB_NJ; ++i) n.config[i] = q[i];
  if (ma) {
    if (t.start_tree && near_hand_index < ma->num_points - 2)
      n.cart_hand_pose_index = near_hand_index + 1;
    else if (!t.start_tree && near_hand_index > 0)
      n.cart_hand_pose_index = near_hand_index - 1;
    else
      n.cart_hand_pose_index = near_hand_index;
  }
  t.nodes.push_back(n);
}

// RRT_Planner::extendTree (RP:787-884).
inline int extend_tree(const PlanParams& pp, Tree& t, const double* q_rand, const double* w, double step, int* q_near_idx,
                       const MaConstraint* ma = nullptr) {
  int nn = nearest_in_tree(t, q_rand);
  *q_near_idx = nn;
  double q_new[NWB_NJ], q_mod[NWB_NJ], q_manip[NWB_NJ];
  int state = generate_q_new(t.nodes[nn].config, q_rand, w, step, q_new);
  int ds = enforce_ds_constraints(pp, q_new, q_mod);
  if (ds == NWB_TRAPPED) return NWB_TRAPPED;
  if (ds == NWB_ADVANCED) state = NWB_ADVANCED;
  if (ma) {
    int mas = enforce_ma_constraints(*ma, t.start_tree, t.nodes[nn].cart_hand_pose_index, t.nodes[nn].config, q_mod, q_manip);
    if (mas == NWB_TRAPPED) return NWB_TRAPPED;
    if (mas == NWB_ADVANCED) state = NWB_ADVANCED;
  } else {
    for (int i = 0; i < NWB_NJ; ++i) q_manip[i] = q_mod[i];
  }
  if (!is_config_valid(pp, q_manip)) return NWB_TRAPPED;
  add_config(t, t.nodes[nn].index, q_manip, ma, t.nodes[nn].cart_hand_pose_index);
  return state;
}

// RRT_Planner::connectTree (RP:890-1034).  *q_near_idx is the node that finally connected
// (RP:1000-1001) or the initial nearest neighbour.  connect_hand_index = q_connect.cart_hand_pose_index.
inline int connect_tree(const PlanParams& pp, Tree& t, const double* q_connect, const double* w, double step,
                        int* q_near_idx, int* steps_out = nullptr, const MaConstraint* ma = nullptr,
                        int connect_hand_index = 0) {
  int nn = nearest_in_tree(t, q_connect);
  *q_near_idx = nn;
  int cur = nn;
  if (steps_out) *steps_out = 0;
  if (ma && t.start_tree && t.nodes[nn].cart_hand_pose_index > connect_hand_index) return NWB_TRAPPED;   // RP:917
  if (ma && !t.start_tree && t.nodes[nn].cart_hand_pose_index < connect_hand_index) return NWB_TRAPPED;  // RP:921
  int state = NWB_ADVANCED, steps = 0;
  while (state != NWB_TRAPPED && state != NWB_REACHED) {
    double q_new[NWB_NJ], q_mod[NWB_NJ], q_manip[NWB_NJ];
    state = generate_q_new(t.nodes[cur].config, q_connect, w, step, q_new);
    ++steps;
    int ds = enforce_ds_constraints(pp, q_new, q_mod);
    if (ds == NWB_TRAPPED) {
      state = NWB_TRAPPED;
      break;
    }
    if (ds == NWB_ADVANCED) state = NWB_ADVANCED;
    bool proceed = true;
    if (ma) {
      int mas = enforce_ma_constraints(*ma, t.start_tree, t.nodes[cur].cart_hand_pose_index, t.nodes[cur].config, q_mod, q_manip);
      if (mas == NWB_TRAPPED) proceed = false;
      if (mas == NWB_ADVANCED) state = NWB_ADVANCED;
      if (t.nodes[cur].cart_hand_pose_index == connect_hand_index && state != NWB_REACHED) proceed = false;  // RP:968
    } else {
      for (int i = 0; i < NWB_NJ; ++i) q_manip[i] = q_mod[i];
    }
    if (!proceed || !is_config_valid(pp, q_manip)) {
      state = NWB_TRAPPED;
      break;
    }
    add_config(t, t.nodes[cur].index, q_manip, ma, t.nodes[cur].cart_hand_pose_index);
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
// (RP:227-302, RP:1039-1212, RP:1303-1392); `ma` = the active articulation constraint or nullptr.
inline SolveResult solve_query(const PlanParams& pp, const double* start, const double* goal, const double* db,
                               int n_db, const double* w, uint64_t seed, int max_iter, double step,
                               const MaConstraint* ma = nullptr) {
  SolveResult res;
  if (!is_config_valid(pp, start) || !is_config_valid(pp, goal)) return res;
  Tree Ts, Tg;
  Tg.start_tree = false;
  add_config(Ts, 0, start);
  add_config(Tg, 0, goal);
  if (ma) {  // RP:274-282
    Ts.nodes[0].cart_hand_pose_index = 0;
    Tg.nodes[0].cart_hand_pose_index = ma->num_points - 1;
  }
  bool swap = false;
  for (int it = 0; it < max_iter; ++it) {
    res.iterations = it + 1;
    const double* q_rand = db + (size_t)random_db_index(pp.quirks, seed, (uint64_t)it, n_db) * NWB_NJ;
    Tree& Ta = swap ? Tg : Ts;
    Tree& Tb = swap ? Ts : Tg;
    int q_near = -1;
    int returned = extend_tree(pp, Ta, q_rand, w, step, &q_near, ma);
    if (returned != NWB_TRAPPED) {
      int found = connect_tree(pp, Tb, Ta.nodes.back().config, w, step, &q_near, nullptr, ma,
                               Ta.nodes.back().cart_hand_pose_index);
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

// ---------------------------------------------------------------------------------------------
// Goal-configuration sampling.

// ConstraintKinematics::compute_ergonomics_index (NCK:612-659): q0 * |"determinant"| where the
// "determinant" of the 6x5 arm Jacobian is the shipped diagonal-product formula over [J J]:
// forward sum of the 5 wrapped diagonals minus a "backward" sum that multiplies ONE column down all
// six rows (quirk C10: not an anti-diagonal).
inline double ergonomics_index(const double* arm5) {
  double J[6 * 5];
  kdlr::jacobian(chains().right_arm, arm5, J);
  const int cols = 5, nc = 10;
  double M[6][10];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < cols; ++j) {
      M[i][j] = J[i * cols + j];
      M[i][j + cols] = M[i][j];
    }
  double sum_fw = 0.0, sum_bw = 0.0;
  for (int i = 0; i < cols; ++i) {
    sum_fw = sum_fw + (M[0][i] * M[1][i + 1] * M[2][i + 2] * M[3][i + 3] * M[4][i + 4] * M[5][i + 5]);
    int c = nc - (1 + i);
    sum_bw = sum_bw - (M[0][c] * M[1][c] * M[2][c] * M[3][c] * M[4][c] * M[5][c]);
  }
  double det_jac = std::fabs(sum_fw - sum_bw);
  return arm5[0] * det_jac;
}

// ConstraintKinematics::computeRightArmIK (NCK:224-506).
// Generated-solver branch (NCK:374-506): among the solutions inside the strict limits keep the one with the
// LARGEST ergonomics index, starting from 0.0 - a posture with RShoulderPitch <= 0 can therefore never be
// chosen (quirk C15).
// Position-only branch (NCK:291-372), taken when the requested direction is exactly (0,0,0): KDL NR on the
// right-arm chain towards Frame(position in the HIP frame) - identity rotation, so the hand orientation is
// constrained after all - from up to 5 random seeds inside the arm limits; the ergonomics index is left
// untouched.  The seeds come from stream kStreamArmSeed of the sample (slot 5 * attempt + joint).
inline bool goal_right_arm_ik(const PlanParams& pp, const Frame& hip, const double* pose6, const double* mn5,
                              const double* mx5, uint64_t seed, uint64_t sample_index, double* out5, double* ergonomics) {
  if (pose6[3] == 0.0 && pose6[4] == 0.0 && pose6[5] == 0.0) {
    kdlr::Rot R = kdlr::transpose(hip.M);
    Frame target;
    double ph[3];
    for (int i = 0; i < 3; ++i) {
      double tmp = -R(i, 0) * hip.p.x - R(i, 1) * hip.p.y - R(i, 2) * hip.p.z;
      double tmp2 = R(i, 0) * pose6[0] + R(i, 1) * pose6[1] + R(i, 2) * pose6[2];
      ph[i] = tmp + tmp2;
    }
    target.p = {ph[0], ph[1], ph[2]};
    int rc = -3;
    bool valid = false;
    double q_init[5], q_out[5];
    for (int iter = 0; iter < 5; ++iter) {                                // NCK:329 max_ik_iter = 5
      for (int i = 0; i < 5; ++i)
        q_init[i] = (mx5[i] - mn5[i]) * stream_uniform(seed, kStreamArmSeed, sample_index, (uint32_t)(5 * iter + i)) + mn5[i];
      rc = kdlr::ik_pos_nr(chains().right_arm, q_init, target, q_out, pp.ik_maxiter, pp.ik_eps, pp.pinv_eps);
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
  double pos[3], dir[3], sols[8 * 5];
  hand_target_in_torso(hip, pose6, pos, dir);
  int n = arm5d_ik(pos, dir, sols);
  if (n == 0) return false;
  double best = 0.0;
  int best_i = 1000;
  for (int s = 0; s < n; ++s) {
    if (!check_joint_limits(sols + s * 5, mn5, mx5, 5)) continue;
    double e = ergonomics_index(sols + s * 5);
    if (e > best) {
      best = e;
      best_i = s;
    }
  }
  if (best_i == 1000) return false;
  *ergonomics = best;
  for (int i = 0; i < 5; ++i) out5[i] = sols[best_i * 5 + i];
  return true;
}

constexpr double kGoalLeftArm[5] = {1.5708, 0.17, 0.0, -0.036, 0.0};  // SCG:296-300
constexpr int kGoalSamplesWanted = 6;                                  // SCG:429 `goal_samples > 5`

// One sample of the goal branch of sample_stable_configs (SCG:278-436): returns 0 = rejected,
// 1 = both IKs solved but infeasible, 2 = accepted (q, ergonomics valid).
inline int goal_sample(const PlanParams& pp, uint64_t seed, uint64_t i, const double* pose6, int max_ik_iter, double* q,
                       double* ergonomics) {
  for (int j = 0; j < NWB_NJ; ++j) {
    double u = stream_uniform(seed, kStreamSampler, i, (uint32_t)j);
    q[j] = (nwbm::kJointMax[j] - nwbm::kJointMin[j]) * u + nwbm::kJointMin[j];
  }
  q[15] = 0.0;
  for (int k = 0; k < 5; ++k) q[5 + k] = kGoalLeftArm[k];
  double rleg[5], lleg[5], rarm[5];
  for (int k = 0; k < 5; ++k) rleg[k] = -q[k];
  if (!left_leg_ik_sampler(pp, rleg, nwbm::kJointMin + 16, nwbm::kJointMax + 16, max_ik_iter, lleg)) return 0;
  Frame hip = kdlr::fk(chains().right_leg, rleg);
  if (!goal_right_arm_ik(pp, hip, pose6, nwbm::kJointMin + 10, nwbm::kJointMax + 10, seed, i, rarm, ergonomics)) return 0;
  for (int k = 0; k < 5; ++k) q[16 + k] = lleg[k];
  for (int k = 0; k < 5; ++k) q[10 + k] = rarm[k];
  return is_feasible(pp.eval, q, nullptr, nullptr) ? 2 : 1;
}

// "%g" text round trip of the goal files (SCG:409-417 written, SCG:540-560 / 666-700 re-read).
inline double through_text(double v) {
  char buf[64];
  std::snprintf(buf, sizeof buf, "%g", v);
  return std::strtod(buf, nullptr);
}

struct GoalResult {
  int num_found = 0;              // return value of sample_stable_configs (<= 6)
  uint64_t samples_drawn = 0;     // loop iterations consumed
  std::vector<double> rows;       // [num_found][21] sorted by ergonomics (descending), text-rounded
  std::vector<double> ergonomics; // same order
  std::vector<uint64_t> index;    // sample index of each row
};

// sample_stable_configs(max_samples, max_ik_iter, t, true, file) + sort_goal_configs_by_ergonomy
// (SCG:176-507, 510-661); loadGoalConfig (SCG:666-756) then takes rows[0].
inline GoalResult sample_goal_configs(const PlanParams& pp, uint64_t seed, const double* pose6, int max_samples,
                                      int max_ik_iter) {
  GoalResult g;
  std::vector<double> rows, erg;
  for (uint64_t i = 0; i < (uint64_t)max_samples; ++i) {
    g.samples_drawn = i + 1;
    double q[NWB_NJ], e = 0.0;
    if (goal_sample(pp, seed, i, pose6, max_ik_iter, q, &e) != 2) continue;
    erg.push_back(through_text(e));
    for (int j = 0; j < NWB_NJ; ++j) rows.push_back(through_text(q[j]));
    g.index.push_back(i);
    if ((int)erg.size() > kGoalSamplesWanted - 1) break;
  }
  g.num_found = (int)erg.size();
  // SCG:585-612: n rounds of "largest remaining value > 0.0", the winner zeroed afterwards
  std::vector<double> e2 = erg;
  std::vector<int> order(erg.size());
  int imax = 0;
  for (size_t n = 0; n < e2.size(); ++n) {
    double vmax = 0.0;
    for (size_t i = 0; i < e2.size(); ++i)
      if (e2[i] > vmax) {
        vmax = e2[i];
        imax = (int)i;
      }
    order[n] = imax;
    e2[imax] = 0.0;
  }
  std::vector<uint64_t> idx2;
  for (size_t n = 0; n < order.size(); ++n) {
    g.rows.insert(g.rows.end(), rows.begin() + order[n] * NWB_NJ, rows.begin() + (order[n] + 1) * NWB_NJ);
    g.ergonomics.push_back(erg[order[n]]);
    idx2.push_back(g.index[order[n]]);
  }
  g.index = idx2;
  return g;
}

// ---------------------------------------------------------------------------------------------
// MoveIt trajectory_processing::IterativeParabolicTimeParameterization::computeTimeStamps as called by
// RRT_Planner::generateTrajectory (RP:1596-1618), followed by the 2 s offset of RP:1614-1618.
// MoveIt is not vendored in the reference; restated from its published algorithm (Hydro/Indigo vintage:
// max_iterations 100, max_time_change_per_it 0.01, rounding threshold 0.01, interval growth factor 1.01).
// PARITY UNPINNED: no shipped file carries time stamps, and the velocity limits come from a URDF that is
// not in the reference tree.
struct TimeParamLimits {
  double v_max[NWB_NJ], a_max[NWB_NJ];
  int max_iterations = 100;
  double max_time_change = 0.01;
};
inline double iptp_grow(double dq1, double dq2, double dt1, double dt2, double a_max, bool grow_first) {
  double a = 2.0 * (dq2 / dt2 - dq1 / dt1) / (dt1 + dt2);
  while (std::fabs(a) > a_max) {
    a = 2.0 * (dq2 / dt2 - dq1 / dt1) / (dt1 + dt2);
    if (grow_first) dt1 *= 1.01;
    else dt2 *= 1.01;
  }
  return grow_first ? dt1 : dt2;
}
inline std::vector<double> time_parameterize(const TimeParamLimits& L, const std::vector<double>& path) {
  const int n = (int)(path.size() / NWB_NJ);
  std::vector<double> stamps(n, 2.0);
  if (n < 2) return stamps;
  auto q = [&](int i, int j) { return path[(size_t)i * NWB_NJ + j]; };
  std::vector<double> dt(n - 1, 0.0);
  for (int i = 0; i + 1 < n; ++i)                                   // applyVelocityConstraints
    for (int j = 0; j < NWB_NJ; ++j) dt[i] = std::max(dt[i], std::fabs(q(i + 1, j) - q(i, j)) / L.v_max[j]);
  int updates, iteration = 0;                                      // applyAccelerationConstraints
  bool backwards = false;
  do {
    updates = 0;
    ++iteration;
    for (int j = 0; j < NWB_NJ; ++j)
      for (int pass = 0; pass < 2; ++pass, backwards = !backwards)
        for (int i = 0; i + 1 < n; ++i) {
          const int k = backwards ? n - 1 - i : i;
          const int prev = k == 0 ? 1 : k - 1, next = k == n - 1 ? n - 2 : k + 1;   // mirrored neighbour at the ends
          const double d1 = q(k, j) - q(prev, j), d2 = q(next, j) - q(k, j);
          const int i1 = k == 0 ? 0 : k - 1, i2 = k == n - 1 ? n - 2 : k;           // intervals before / after point k
          const double t1 = dt[i1], t2 = dt[i2];
          if (t1 == 0.0 || t2 == 0.0) continue;
          const double a = 2.0 * (d2 / t2 - d1 / t1) / (t1 + t2);
          if (std::fabs(a) > L.a_max[j] + 0.01) {
            if (!backwards) dt[i2] = std::min(t2 + L.max_time_change, iptp_grow(d1, d2, t1, t2, L.a_max[j], false));
            else dt[i1] = std::min(t1 + L.max_time_change, iptp_grow(d1, d2, t1, t2, L.a_max[j], true));
            ++updates;
          }
        }
  } while (updates > 0 && iteration < L.max_iterations);
  double t = 0.0;
  for (int i = 1; i < n; ++i) {
    t += dt[i - 1];
    stamps[i] = t + 2.0;
  }
  return stamps;
}

// ---------------------------------------------------------------------------------------------
// Service level (NPC:150-870).  Wall-clock outputs (time_goal_config, search_time) are not part of
// the restatement.  Seeds: goal sampling uses `seed` (approach) and `seed + 1` (manipulation),
// planning uses `seed + 2` and `seed + 3` - a convention shared with the product.
struct ServiceParams {                 // NPC:20-33 defaults overridden by planning_parameters.yaml
  int max_samples = 5000, max_ik_iterations = 5, max_expand_iterations = 8000;
  double advance_steps = 0.1;
  double w_approach[3] = {1.0, 1.0, 1.0};      // r_arm, l_arm, legs
  double w_manipulate[3] = {1.0, 0.0, 1.0};
  int shortcut_waypoints = 100, manipulate_waypoints = 4, shortcut_rounds = 500;   // RP:1397-1398, RP:1851
  int num_interp_points = 20;                  // NPC:539,822
  bool time_parameterization = true;           // RP:1569-1618
  TimeParamLimits limits;
  ServiceParams() {
    // NAO V4 URDF velocity limits in planner joint order, from memory (UNVERIFIED, see include/nwb.h); no
    // acceleration limits in the URDF -> MoveIt's default 1.0
    const double v[NWB_NJ] = {4.16174, 6.40239, 6.40239, 6.40239, 4.16174, 8.26797, 7.19407, 8.26797, 7.19407, 24.6229,
                              8.26797, 7.19407, 8.26797, 7.19407, 24.6229, 4.16174, 4.16174, 6.40239, 6.40239, 6.40239, 4.16174};
    for (int j = 0; j < NWB_NJ; ++j) {
      limits.v_max[j] = v[j];
      limits.a_max[j] = 1.0;
    }
  }
};
inline void limb_weights(const double* w3, double* w21) {  // NPC:280-306
  for (int i = 0; i < 5; ++i) w21[i] = w3[2];
  for (int i = 5; i < 10; ++i) w21[i] = w3[1];
  for (int i = 10; i < 15; ++i) w21[i] = w3[0];
  w21[15] = 0.0;
  for (int i = 16; i < 21; ++i) w21[i] = w3[2];
}

struct PlanOut {
  bool success = false;
  int num_nodes = 0;
  std::vector<double> trajectory;      // way-points of the DisplayTrajectory
  std::vector<double> time_from_start; // before time parameterisation: 2.0 + i * increment (RP:1532-1553)
};

// writePath after a successful solveQuery (RP:1303-1455): shortcut + densify (free motion) or
// densify with 4 way-points (articulated motion); `execute` = setTrajectoryMode(true).
inline void finish_path(const PlanParams& pp, const ServiceParams& sp, const SolveResult& r, bool articulated,
                        bool execute, double time_step_manip, PlanOut& out) {
  std::vector<double> path;
  if (!articulated) {
    std::vector<double> shortp;
    bool ok = path_shortcutter(pp, r.raw_path, sp.shortcut_waypoints, sp.shortcut_rounds, shortp);
    if (ok) {
      if (!execute) interpolate_short_path(shortp, sp.shortcut_waypoints, path);
      else path = shortp;
    }   // RP:1919-1926: on failure shortcutted_path stays empty
  } else {
    if (!execute) interpolate_short_path(r.raw_path, sp.manipulate_waypoints, path);
    else path = r.raw_path;
  }
  out.trajectory = path;
  float increment = 2.0f;
  for (size_t i = 0; i < path.size() / NWB_NJ; ++i) {
    out.time_from_start.push_back((double)increment);
    increment = articulated ? increment + (float)time_step_manip : increment + 1.0f;
  }
  if (sp.time_parameterization && !(execute && articulated) && !path.empty()) out.time_from_start = time_parameterize(sp.limits, path);
  out.success = !path.empty();
  out.num_nodes = r.num_nodes;
}

struct MotionPlanOut {
  GoalResult goal;
  PlanOut plan;
};
// NAO_PLANNER_CONTROL::compute_motion_plan (NPC:183-336).
inline MotionPlanOut compute_motion_plan(const PlanParams& pp, const ServiceParams& sp, const double* db, int n_db,
                                         const double* pose6, const double* start, bool execute, uint64_t seed) {
  MotionPlanOut o;
  o.goal = sample_goal_configs(pp, seed, pose6, sp.max_samples, sp.max_ik_iterations);
  if (o.goal.num_found == 0) return o;
  double w[NWB_NJ];
  limb_weights(sp.w_approach, w);
  SolveResult r = solve_query(pp, start, o.goal.rows.data(), db, n_db, w, seed + 2, sp.max_expand_iterations, sp.advance_steps);
  if (r.success) finish_path(pp, sp, r, false, execute, 0.0, o.plan);
  return o;
}

struct ManipulationPlanOut {
  GoalResult goal_approach, goal_manipulate;
  PlanOut approach, manipulation;
  double pose_manipulate[6] = {0, 0, 0, 0, 0, 0};
};
// compute_linear_manipulation_plan (NPC:338-588) and compute_circular_manipulation_plan (NPC:590-870):
// `object` = {relative_angle, translation_length} (linear) or {relative_angle, rotation_radius,
// rotation_angle, clearance_handle} (circular).
inline ManipulationPlanOut compute_manipulation_plan(const PlanParams& pp, const ServiceParams& sp, const double* db,
                                                     int n_db, const double* pose6, const double* start, bool circular,
                                                     const double* object, bool execute, uint64_t seed) {
  ManipulationPlanOut o;
  o.goal_approach = sample_goal_configs(pp, seed, pose6, sp.max_samples, sp.max_ik_iterations);
  if (o.goal_approach.num_found == 0) return o;
  double w[NWB_NJ];
  limb_weights(sp.w_approach, w);
  const double* goal_a = o.goal_approach.rows.data();
  SolveResult r = solve_query(pp, start, goal_a, db, n_db, w, seed + 2, sp.max_expand_iterations, sp.advance_steps);
  if (r.success) finish_path(pp, sp, r, false, execute, 0.0, o.approach);
  if (!o.approach.success) return o;
  double* pm = o.pose_manipulate;
  MaConstraint ma;
  float time_step;
  const float manipulation_velocity = 0.02f;   // RP:92
  if (!circular) {
    double pos_angle = (90.0 - object[0]) * (M_PI / 180);                          // NPC:502
    pm[0] = pose6[0] - object[1] * std::sin(pos_angle);
    pm[1] = pose6[1] + object[1] * std::cos(pos_angle);
    pm[2] = pose6[2];
  } else {                                                                          // NPC:762-790
    double rel = object[0] * (M_PI / 180);
    double x_door = pose6[0] + object[3] * std::cos(rel);
    double y_door = pose6[1] - object[3] * std::sin(rel);
    double z_door = pose6[2];
    double x_hinge = x_door - object[1] * std::sin(rel);
    double y_hinge = y_door - object[1] * std::cos(rel);
    double tmp_angle = (object[2] - object[0]) * (M_PI / 180);
    x_door = x_hinge - object[1] * std::sin(tmp_angle);
    y_door = y_hinge + object[1] * std::cos(tmp_angle);
    tmp_angle = 1.5708 - tmp_angle;
    pm[0] = x_door - object[3] * std::sin(tmp_angle);
    pm[1] = y_door - object[3] * std::cos(tmp_angle);
    pm[2] = z_door;
  }
  for (int i = 3; i < 6; ++i) pm[i] = pose6[i];
  o.goal_manipulate = sample_goal_configs(pp, seed + 1, pm, sp.max_samples, sp.max_ik_iterations);
  if (o.goal_manipulate.num_found == 0) return o;
  if (!circular) {                                                                  // RP:195-206
    ma = drawer_constraint(pose6, pm, sp.num_interp_points);
    float manipulation_time = (float)(object[1]) / manipulation_velocity;
    time_step = manipulation_time / float(sp.num_interp_points + 2);
  } else {                                                                          // NPC:816-822, RP:209-223
    double door[4] = {object[1], object[2], object[0], object[3]};
    ma = door_constraint(pose6, pm, sp.num_interp_points, door);
    float arc_length = float(door[0] * M_PI * door[1]) / 180.0f;
    float manipulation_time = arc_length / manipulation_velocity;
    time_step = manipulation_time / float(sp.num_interp_points + 2);
  }
  limb_weights(sp.w_manipulate, w);
  SolveResult r2 = solve_query(pp, goal_a, o.goal_manipulate.rows.data(), db, n_db, w, seed + 3, sp.max_expand_iterations,
                               sp.advance_steps, &ma);
  if (r2.success) finish_path(pp, sp, r2, true, execute, (double)time_step, o.manipulation);
  return o;
}

}  // namespace oracle
