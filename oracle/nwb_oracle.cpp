// TEST INFRASTRUCTURE - CPU oracle, never shipped, never on the product path.
//
// extern "C" surface of the fp64 restatement (oracle/kdl_restated.h, model_eval.h,
// planner_restated.h), loaded through ctypes by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs ONLY.  The product library (libnwb.so) never links,
// loads or calls anything in here.
//
// Parity status: the reference cannot be compiled in this environment (it needs ROS/catkin,
// MoveIt, FCL, orocos-KDL, hrl_kinematics and the nao_description2 URDF, none vendored - see
// DESIGN.md), so this restatement is pinned against the reference's shipped data files:
// KDL-chain FK and NR/pinv IK two-sidedly, the support hull
// against the logged polygon (E10), collision + stability ONE-SIDEDLY (E12: everything the
// reference shipped as valid stays valid).  Collision geometry and mass distribution are
// authored => for those two inputs: PARITY UNPINNED beyond the one-sided gate.
#include <atomic>
#include <cstring>
#include <thread>

#include "arm_restated.h"
#include "planner_restated.h"

using namespace oracle;

namespace {
PlanParams make_params(const nwb_params* p, const double* boxes, int nb) {
  PlanParams pp;
  if (p) {
    pp.eval.scale_sp = p->scale_sp;
    pp.eval.scale_origin = p->scale_origin;
    pp.eval.srdf_trim = p->srdf_trim;
    pp.quirks = (unsigned)p->quirks;
    pp.ds_tolerance = p->ds_tolerance;
    pp.ik_eps = p->ik_eps;
    pp.ik_maxiter = p->ik_maxiter;
    pp.pinv_eps = p->pinv_eps;
  }
  for (int i = 0; i < nb; ++i) {
    const double* b = boxes + (size_t)i * 15;
    Box bx;
    bx.c = {b[0], b[1], b[2]};
    bx.h = {b[3] * 0.5, b[4] * 0.5, b[5] * 0.5};
    for (int k = 0; k < 9; ++k) bx.R.m[k] = b[6 + k];
    pp.eval.boxes.push_back(bx);
  }
  return pp;
}
void put_frame(const Frame& f, double* o) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) o[4 * r + c] = f.M(r, c);
    o[4 * r + 3] = r == 0 ? f.p.x : r == 1 ? f.p.y : f.p.z;
  }
}
template <class F>
void parallel_for(int64_t n, int threads, F fn) {
  if (threads <= 1) {
    fn(0, n);
    return;
  }
  std::vector<std::thread> th;
  int64_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    int64_t a = t * chunk, b = std::min(n, a + chunk);
    if (a >= b) break;
    th.emplace_back([=] { fn(a, b); });
  }
  for (auto& t : th) t.join();
}
}  // namespace

extern "C" {

int nwbo_num_frames() { return nwbm::kNumFrames; }
int nwbo_num_pairs(int srdf_trim, int group_only) {
  if (srdf_trim) return group_only ? nwbm::kNumPairsGroupTrim : nwbm::kNumPairsTrim;
  return group_only ? nwbm::kNumPairsGroupNoTrim : nwbm::kNumPairsNoTrim;
}
int nwbo_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

void nwbo_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
double nwbo_stream_uniform(uint64_t seed, uint32_t stream, uint64_t index, uint32_t slot) {
  return stream_uniform(seed, stream, index, slot);
}

// MoveIt RobotState FK + hrl_kinematics computeCOM (called rrt_planner.cpp:715-720,1275).
void nwbo_fk_batch(const double* q, int64_t n, double* poses /*[n][F][12]*/, double* com /*[n][3]*/) {
  for (int64_t i = 0; i < n; ++i) {
    Frame fr[nwbm::kNumFrames];
    fk_all(q + i * NWB_NJ, fr);
    if (poses)
      for (int f = 0; f < nwbm::kNumFrames; ++f) put_frame(fr[f], poses + ((size_t)i * nwbm::kNumFrames + f) * 12);
    if (com) {
      Vec c = center_of_mass(fr);
      com[i * 3 + 0] = c.x; com[i * 3 + 1] = c.y; com[i * 3 + 2] = c.z;
    }
  }
}

// RRT_Planner::isFeasible (rrt_planner.cpp:1219-1287).
void nwbo_is_feasible_batch(const nwb_params* p, const double* boxes, int nb, const double* q, int64_t n,
                            uint8_t* valid, uint8_t* reason, double* margins /*[n][2]*/, int threads) {
  PlanParams pp = make_params(p, boxes, nb);
  parallel_for(n, threads, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      unsigned r;
      double m[2];
      bool ok = is_feasible(pp.eval, q + i * NWB_NJ, &r, m);
      if (valid) valid[i] = ok;
      if (reason) reason[i] = (uint8_t)r;
      if (margins) {
        margins[2 * i] = m[0];
        margins[2 * i + 1] = m[1];
      }
    }
  });
}

// isStateColliding over the whole robot (the per-way-point test of isPathValid, rrt_planner.cpp:1876).
void nwbo_is_state_colliding_batch(const nwb_params* p, const double* boxes, int nb, const double* q, int64_t n,
                                   uint8_t* colliding, double* min_sep, int group_only) {
  PlanParams pp = make_params(p, boxes, nb);
  for (int64_t i = 0; i < n; ++i) {
    Frame fr[nwbm::kNumFrames];
    fk_all(q + i * NWB_NJ, fr);
    double s;
    bool c = in_collision(pp.eval, fr, group_only != 0, &s);
    if (colliding) colliding[i] = c;
    if (min_sep) min_sep[i] = s;
  }
}

// Support hull of TestStability::isPoseStable for one configuration (clockwise, no closing point).
int nwbo_support_hull(const nwb_params* p, const double* q, double* xy /*[cap][2]*/, int cap) {
  PlanParams pp = make_params(p, nullptr, 0);
  Frame fr[nwbm::kNumFrames];
  fk_all(q, fr);
  std::vector<P2> h = support_hull(pp.eval, fr);
  int n = std::min((int)h.size(), cap);
  for (int i = 0; i < n; ++i) {
    xy[2 * i] = h[i].x;
    xy[2 * i + 1] = h[i].y;
  }
  return (int)h.size();
}

double nwbo_seg_seg_distance(const double* p1, const double* q1, const double* p2, const double* q2) {
  return seg_seg_distance({p1[0], p1[1], p1[2]}, {q1[0], q1[1], q1[2]}, {p2[0], p2[1], p2[2]}, {q2[0], q2[1], q2[2]});
}
double nwbo_seg_box_distance(const double* a, const double* b, const double* box15) {
  PlanParams pp = make_params(nullptr, box15, 1);
  return seg_box_distance({a[0], a[1], a[2]}, {b[0], b[1], b[2]}, pp.eval.boxes[0]);
}

// KDL chain FK: which = 0 right leg (5), 1 left leg (5), 2 legs (10), 3 right arm (5).
void nwbo_chain_fk(int which, const double* q, double* frame12) {
  const Chains& c = chains();
  const kdlr::Chain& ch = which == 0 ? c.right_leg : which == 1 ? c.left_leg : which == 2 ? c.legs : c.right_arm;
  put_frame(kdlr::fk(ch, q), frame12);
}
// right-hand grasp frame in the r_sole frame: hip(right_leg_values) * arm(q[10..14]) (NCK:303-334).
void nwbo_right_hand_pose(const double* q21, double* frame12) {
  double rl[5];
  for (int i = 0; i < 5; ++i) rl[i] = -q21[i];
  Frame hip = kdlr::fk(chains().right_leg, rl);
  put_frame(hip * kdlr::fk(chains().right_arm, q21 + 10), frame12);
}
void nwbo_chain_jacobian(int which, const double* q, double* J) {
  const Chains& c = chains();
  const kdlr::Chain& ch = which == 0 ? c.right_leg : which == 1 ? c.left_leg : which == 2 ? c.legs : c.right_arm;
  kdlr::jacobian(ch, q, J);
}

// ChainIkSolverVel_pinv on a chain at q (twist 6 -> qdot n) and the same pseudo-inverse application on an arbitrary
// row-major 6 x n matrix (tests/test_pinv_svd_cpu.py checks both against numpy.linalg.svd, incl. sigma_min near eps)
void nwbo_ik_vel_pinv(int which, const double* q, const double* twist, double eps, double* qdot) {
  const Chains& c = chains();
  const kdlr::Chain& ch = which == 0 ? c.right_leg : which == 1 ? c.left_leg : which == 2 ? c.legs : c.right_arm;
  kdlr::ik_vel_pinv(ch, q, twist, qdot, eps);
}
void nwbo_pinv_apply(const double* J, int n, const double* twist, double eps, double* qdot, double* S_out) {
  std::vector<double> U(6 * n), S(n), V(n * n), tmp(n);
  kdlr::svd_jacobi(J, 6, n, U.data(), S.data(), V.data());
  for (int i = 0; i < n; ++i) {
    double sum = 0;
    for (int j = 0; j < 6; ++j) sum += U[j * n + i] * twist[j];
    tmp[i] = sum * (std::fabs(S[i]) < eps ? 0.0 : 1.0 / S[i]);
    if (S_out) S_out[i] = S[i];
  }
  for (int i = 0; i < n; ++i) {
    double sum = 0;
    for (int j = 0; j < n; ++j) sum += V[i * n + j] * tmp[j];
    qdot[i] = sum;
  }
}

int nwbo_check_joint_limits(const double* q21) { return check_joint_limits(q21, nwbm::kJointMin, nwbm::kJointMax, NWB_NJ); }
int nwbo_check_left_leg_pose(const nwb_params* p, const double* legs10) {
  return check_left_leg_pose(make_params(p, nullptr, 0), legs10);
}
int nwbo_left_leg_ik_local(const nwb_params* p, const double* legs10, double* out5, int* iters) {
  return left_leg_ik_local(make_params(p, nullptr, 0), legs10, nwbm::kJointMin + 16, nwbm::kJointMax + 16, out5, iters);
}
int nwbo_left_leg_ik_sampler(const nwb_params* p, const double* right_leg_values5, int max_ik_iter, double* out5,
                             int* iters) {
  return left_leg_ik_sampler(make_params(p, nullptr, 0), right_leg_values5, nwbm::kJointMin + 16, nwbm::kJointMax + 16,
                             max_ik_iter, out5, iters);
}

// RRT_Planner::findNearestNeighbour for nq queries over the same node array.
// as_shipped != 0 additionally deep-copies the node array per query, like the by-value Tree
// parameter of rrt_planner.h:160 (timing only; results identical).
void nwbo_nearest_neighbour_batch(const double* nodes, int64_t N, const double* q, int64_t nq, int32_t* idx,
                                  double* dist, int as_shipped, int threads) {
  parallel_for(nq, threads, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      double d;
      if (as_shipped) {
        std::vector<std::vector<double>> copy((size_t)N);
        for (int64_t k = 0; k < N; ++k) copy[k].assign(nodes + k * NWB_NJ, nodes + (k + 1) * NWB_NJ);
        int nn = -1;
        double best = 10000.0;
        for (int64_t k = 0; k < N; ++k) {
          double sum = 0;
          for (int j = 0; j < NWB_NJ; ++j) {
            double dd = q[i * NWB_NJ + j] - copy[k][j];
            sum = sum + dd * dd;
          }
          double dn = std::sqrt(sum);
          if (dn < best) {
            best = dn;
            nn = (int)k;
          }
        }
        idx[i] = nn;
        d = best;
      } else {
        idx[i] = nearest_neighbour(nodes, N, q + i * NWB_NJ, &d);
      }
      if (dist) dist[i] = d;
    }
  });
}

void nwbo_generate_q_new_batch(const double* q_near, const double* q_rand, int64_t n, const double* w, double step,
                               double* q_new, int32_t* status) {
  for (int64_t i = 0; i < n; ++i)
    status[i] = generate_q_new(q_near + i * NWB_NJ, q_rand + i * NWB_NJ, w, step, q_new + i * NWB_NJ);
}

// generate_q_new -> enforce_DS_Constraints, combined as extendTree does (rrt_planner.cpp:807-818).
// q_out is written only when status != TRAPPED.  ds_status (optional) = enforce_DS result.
void nwbo_steer_project_batch(const nwb_params* p, const double* q_near, const double* q_target, int64_t n,
                              const double* w, double step, double* q_out, int32_t* status, int32_t* ds_status,
                              int threads) {
  PlanParams pp = make_params(p, nullptr, 0);
  parallel_for(n, threads, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      double q_new[NWB_NJ], q_mod[NWB_NJ];
      int st = generate_q_new(q_near + i * NWB_NJ, q_target + i * NWB_NJ, w, step, q_new);
      int ds = enforce_ds_constraints(pp, q_new, q_mod);
      if (ds_status) ds_status[i] = ds;
      if (ds == NWB_TRAPPED) {
        status[i] = NWB_TRAPPED;
        continue;
      }
      if (ds == NWB_ADVANCED) st = NWB_ADVANCED;
      status[i] = st;
      std::memcpy(q_out + i * NWB_NJ, q_mod, sizeof(q_mod));
    }
  });
}

int nwbo_enforce_ds(const nwb_params* p, const double* q_new, double* q_mod) {
  return enforce_ds_constraints(make_params(p, nullptr, 0), q_new, q_mod);
}

// connectTree walk from q_start towards q_connect as one edge (rrt_planner.cpp:926-1031):
// returns the final status; nodes_out receives the projected nodes that were added (<= cap).
int nwbo_connect_edge(const nwb_params* p, const double* boxes, int nb, const double* q_start, const double* q_connect,
                      const double* w, double step, double* nodes_out, int cap, int* n_added) {
  PlanParams pp = make_params(p, boxes, nb);
  Tree t;
  add_config(t, 0, q_start);
  int near_idx;
  int st = connect_tree(pp, t, q_connect, w, step, &near_idx);
  int added = (int)t.nodes.size() - 1;
  for (int i = 0; i < std::min(added, cap); ++i) std::memcpy(nodes_out + (size_t)i * NWB_NJ, t.nodes[i + 1].config, sizeof(double) * NWB_NJ);
  *n_added = added;
  return st;
}

void nwbo_interpolate_waypoints(const double* a, const double* b, int k, double* out /*[k+2][21]*/) {
  std::vector<double> w;
  interpolate_waypoints(a, b, k, w);
  std::memcpy(out, w.data(), w.size() * sizeof(double));
}

void nwbo_is_path_valid_batch(const nwb_params* p, const double* boxes, int nb, const double* qa, const double* qb,
                              int64_t n_edges, int k, uint8_t* ok, int32_t* first_bad, int threads) {
  PlanParams pp = make_params(p, boxes, nb);
  parallel_for(n_edges, threads, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      int fb;
      bool v = is_path_valid(pp, qa + i * NWB_NJ, qb + i * NWB_NJ, k, &fb);
      ok[i] = v;
      if (first_bad) first_bad[i] = fb;
    }
  });
}

// path_shortcutter way-point selection; returns number of way-points written (<= cap), -1 on budget exhaustion.
int nwbo_shortcut_path(const nwb_params* p, const double* boxes, int nb, const double* raw, int n_raw, double* out,
                       int cap, int* edges_checked) {
  PlanParams pp = make_params(p, boxes, nb);
  std::vector<double> r(raw, raw + (size_t)n_raw * NWB_NJ), o;
  bool ok = path_shortcutter(pp, r, p ? p->shortcut_waypoints : 100, p ? p->shortcut_max_rounds : 500, o, edges_checked);
  int n = (int)(o.size() / NWB_NJ);
  std::memcpy(out, o.data(), sizeof(double) * NWB_NJ * std::min(n, cap));
  return ok ? n : -1;
}

int nwbo_interpolate_short_path(const double* path, int n_path, int k, double* out, int cap) {
  std::vector<double> pth(path, path + (size_t)n_path * NWB_NJ), o;
  interpolate_short_path(pth, k, o);
  int n = (int)(o.size() / NWB_NJ);
  std::memcpy(out, o.data(), sizeof(double) * NWB_NJ * std::min(n, cap));
  return n;
}

// sample_stable_configs database branch over sample indices [first, first+count).
int64_t nwbo_sample_stable_configs(const nwb_params* p, const double* boxes, int nb, uint64_t seed, uint64_t first,
                                   uint64_t count, int max_ik_iter, double* rows_out, int64_t cap,
                                   uint64_t* index_out, uint8_t* stage_out, int threads) {
  PlanParams pp = make_params(p, boxes, nb);
  int T = std::max(1, threads);
  std::vector<std::vector<double>> rows(T);
  std::vector<std::vector<uint64_t>> idx(T);
  std::vector<std::vector<uint8_t>> stg(T);
  uint64_t chunk = (count + T - 1) / T;
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t) {
    uint64_t a = first + t * chunk, b = std::min(first + count, a + chunk);
    if (a >= b) break;
    th.emplace_back([&, t, a, b] { sample_stable_configs(pp, seed, a, b - a, max_ik_iter, rows[t], idx[t], &stg[t]); });
  }
  for (auto& x : th) x.join();
  int64_t n = 0;
  uint64_t s = 0;
  for (int t = 0; t < T; ++t) {
    for (size_t r = 0; r < idx[t].size(); ++r, ++n) {
      if (n < cap) {
        if (rows_out) std::memcpy(rows_out + n * NWB_NJ, &rows[t][r * NWB_NJ], sizeof(double) * NWB_NJ);
        if (index_out) index_out[n] = idx[t][r];
      }
    }
    if (stage_out)
      for (uint8_t v : stg[t]) stage_out[s++] = v;
  }
  return n;
}

int nwbo_random_db_index(const nwb_params* p, uint64_t seed, uint64_t iteration, int n_db) {
  return random_db_index(p ? (unsigned)p->quirks : NWB_QUIRKS_DEFAULT, seed, iteration, n_db);
}

// setStartGoalConfigs + solveQuery + raw path extraction.  stats = {success, iterations,
// num_nodes, nodes_start, nodes_goal}.  Returns number of raw way-points (0 = no plan).
int nwbo_solve_query(const nwb_params* p, const double* boxes, int nb, const double* start, const double* goal,
                     const double* db, int n_db, const double* w, uint64_t seed, int max_iter, double step,
                     double* raw_out, int cap, int32_t* stats) {
  PlanParams pp = make_params(p, boxes, nb);
  SolveResult r = solve_query(pp, start, goal, db, n_db, w, seed, max_iter, step);
  int n = (int)(r.raw_path.size() / NWB_NJ);
  if (raw_out) std::memcpy(raw_out, r.raw_path.data(), sizeof(double) * NWB_NJ * std::min(n, cap));
  if (stats) {
    stats[0] = r.success; stats[1] = r.iterations; stats[2] = r.num_nodes; stats[3] = r.nodes_start; stats[4] = r.nodes_goal;
  }
  return n;
}

// Right-arm position + direction solver (closed-form restatement of IKF:341-1517, oracle/arm_restated.h).
void nwbo_arm5d_fk(const double* q, double* pos, double* dir) { arm5d_fk(q, pos, dir); }
int nwbo_arm5d_ik(const double* pos, const double* dir, double* out /*[8][5]*/) { return arm5d_ik(pos, dir, out); }

// ---- right arm / articulated-object constraint / goal sampling / services ----
namespace {
MaConstraint make_ma(const double* pts, int num_points) {
  MaConstraint ma;
  ma.num_points = num_points;
  ma.pts.assign(pts, pts + (size_t)num_points * 6);
  return ma;
}
ServiceParams make_service(const double* sv /*[13] or null*/, const nwb_params* p = nullptr) {
  ServiceParams sp;
  if (p) {
    sp.time_parameterization = p->time_parameterization != 0;
    for (int j = 0; j < NWB_NJ; ++j) {
      sp.limits.v_max[j] = p->joint_max_velocity[j];
      sp.limits.a_max[j] = p->joint_max_acceleration[j];
    }
    sp.limits.max_iterations = p->iptp_max_iterations;
    sp.limits.max_time_change = p->iptp_max_time_change;
  }
  if (sv) {
    sp.max_samples = (int)sv[0]; sp.max_ik_iterations = (int)sv[1]; sp.max_expand_iterations = (int)sv[2];
    sp.advance_steps = sv[3];
    for (int i = 0; i < 3; ++i) sp.w_approach[i] = sv[4 + i];
    for (int i = 0; i < 3; ++i) sp.w_manipulate[i] = sv[7 + i];
    sp.shortcut_waypoints = (int)sv[10]; sp.manipulate_waypoints = (int)sv[11]; sp.num_interp_points = (int)sv[12];
  }
  return sp;
}
int put_rows(const std::vector<double>& v, double* out, int cap_rows) {
  int n = (int)(v.size() / NWB_NJ);
  if (out) std::memcpy(out, v.data(), sizeof(double) * NWB_NJ * std::min(n, cap_rows));
  return n;
}
void put_goal(const GoalResult& g, double* rows6, double* erg6, uint64_t* idx6, int32_t* found, uint64_t* drawn) {
  if (rows6) std::memcpy(rows6, g.rows.data(), sizeof(double) * g.rows.size());
  if (erg6) std::memcpy(erg6, g.ergonomics.data(), sizeof(double) * g.ergonomics.size());
  if (idx6) std::memcpy(idx6, g.index.data(), sizeof(uint64_t) * g.index.size());
  if (found) *found = g.num_found;
  if (drawn) *drawn = g.samples_drawn;
}
}  // namespace

// kind 0 = drawer (AOC:105-122), 1 = door (AOC:126-189); out [num_interp + 2][6].
void nwbo_hand_trajectory(int kind, const double* start6, const double* goal6, int num_interp, const double* door4, double* out) {
  MaConstraint ma = kind == 0 ? drawer_constraint(start6, goal6, num_interp) : door_constraint(start6, goal6, num_interp, door4);
  std::memcpy(out, ma.pts.data(), sizeof(double) * ma.pts.size());
}
double nwbo_ergonomics_index(const double* arm5) { return ergonomics_index(arm5); }
// FK of the sole->hand KDL chain (AOC:42-56): position + z-axis of the grasp frame.
void nwbo_right_hand_fk(const double* q21, double* pos, double* zaxis) {
  double q[10];
  for (int i = 0; i < 5; ++i) q[i] = -q21[i];
  for (int i = 0; i < 5; ++i) q[5 + i] = q21[10 + i];
  Frame h = kdlr::fk(chains().r_leg_arm, q);
  pos[0] = h.p.x; pos[1] = h.p.y; pos[2] = h.p.z;
  zaxis[0] = h.M(0, 2); zaxis[1] = h.M(1, 2); zaxis[2] = h.M(2, 2);
}
// computeRightArmIK (NCK:224-506) for whole-body rows: ok[i], arm[i][5], ergonomics[i].
void nwbo_goal_right_arm_ik_batch(const nwb_params* p, const double* q /*[n][21]*/, int64_t n, const double* pose6, uint64_t seed,
                                  uint64_t first_index, uint8_t* ok, double* arm, double* ergonomics) {
  PlanParams pp = make_params(p, nullptr, 0);
  for (int64_t i = 0; i < n; ++i) {
    double rleg[5];
    for (int k = 0; k < 5; ++k) rleg[k] = -q[i * NWB_NJ + k];
    Frame hip = kdlr::fk(chains().right_leg, rleg);
    double e = 0, a[5] = {0, 0, 0, 0, 0};
    ok[i] = goal_right_arm_ik(pp, hip, pose6, nwbm::kJointMin + 10, nwbm::kJointMax + 10, seed, first_index + (uint64_t)i, a, &e);
    for (int k = 0; k < 5; ++k) arm[i * 5 + k] = a[k];
    ergonomics[i] = e;
  }
}
// enforce_MA_Constraints (RP:597-696) for n independent (q_near, q_new_ds) pairs.
void nwbo_enforce_ma_batch(const double* pts, int num_points, int start_tree, const int32_t* near_hand_index,
                           const double* q_near, const double* q_ds, int64_t n, double* q_out, int32_t* status) {
  MaConstraint ma = make_ma(pts, num_points);
  for (int64_t i = 0; i < n; ++i) {
    double o[NWB_NJ];
    for (int j = 0; j < NWB_NJ; ++j) o[j] = std::nan("");
    status[i] = enforce_ma_constraints(ma, start_tree != 0, near_hand_index[i], q_near + i * NWB_NJ, q_ds + i * NWB_NJ, o);
    std::memcpy(q_out + i * NWB_NJ, o, sizeof o);
  }
}
// per-sample outcome of the goal branch for indices [first, first+count): stage 0/1/2, ergonomics, row.
void nwbo_goal_sample_batch(const nwb_params* p, const double* boxes, int nb, uint64_t seed, uint64_t first, uint64_t count,
                            const double* pose6, int max_ik_iter, uint8_t* stage, double* ergonomics, double* rows, int threads) {
  PlanParams pp = make_params(p, boxes, nb);
  parallel_for((int64_t)count, threads, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      double q[NWB_NJ], e = 0;
      stage[i] = (uint8_t)goal_sample(pp, seed, first + (uint64_t)i, pose6, max_ik_iter, q, &e);
      if (ergonomics) ergonomics[i] = e;
      if (rows) std::memcpy(rows + i * NWB_NJ, q, sizeof q);
    }
  });
}
void nwbo_sample_goal_configs(const nwb_params* p, const double* boxes, int nb, uint64_t seed, const double* pose6,
                              int max_samples, int max_ik_iter, double* rows6, double* erg6, uint64_t* idx6, int32_t* found,
                              uint64_t* drawn) {
  PlanParams pp = make_params(p, boxes, nb);
  put_goal(sample_goal_configs(pp, seed, pose6, max_samples, max_ik_iter), rows6, erg6, idx6, found, drawn);
}
// solve_query with an active articulation constraint (pts != null).
int nwbo_solve_query_ma(const nwb_params* p, const double* boxes, int nb, const double* start, const double* goal,
                        const double* db, int n_db, const double* w, uint64_t seed, int max_iter, double step,
                        const double* pts, int num_points, double* raw_out, int cap, int32_t* stats) {
  PlanParams pp = make_params(p, boxes, nb);
  MaConstraint ma;
  if (pts) ma = make_ma(pts, num_points);
  SolveResult r = solve_query(pp, start, goal, db, n_db, w, seed, max_iter, step, pts ? &ma : nullptr);
  int n = put_rows(r.raw_path, raw_out, cap);
  if (stats) {
    stats[0] = r.success; stats[1] = r.iterations; stats[2] = r.num_nodes; stats[3] = r.nodes_start; stats[4] = r.nodes_goal;
  }
  return n;
}
// compute_motion_plan (NPC:183-336).  stats = {success, num_nodes, goal configs found, way-points}.
void nwbo_compute_motion_plan(const nwb_params* p, const double* boxes, int nb, const double* service13, const double* db,
                              int n_db, const double* pose6, const double* start, int execute, uint64_t seed,
                              double* goal_config, double* traj, double* tfs, int cap, int32_t* stats) {
  PlanParams pp = make_params(p, boxes, nb);
  MotionPlanOut o = compute_motion_plan(pp, make_service(service13, p), db, n_db, pose6, start, execute != 0, seed);
  if (o.goal.num_found && goal_config) std::memcpy(goal_config, o.goal.rows.data(), sizeof(double) * NWB_NJ);
  int n = put_rows(o.plan.trajectory, traj, cap);
  if (tfs) std::memcpy(tfs, o.plan.time_from_start.data(), sizeof(double) * std::min(n, cap));
  stats[0] = o.plan.success; stats[1] = o.plan.num_nodes; stats[2] = o.goal.num_found; stats[3] = n;
}
// compute_linear / compute_circular_manipulation_plan (NPC:338-588, 590-870).
// stats = {success_approach, success_manipulation, nodes_approach, nodes_manipulate, goals_approach,
//          goals_manipulate, way-points approach, way-points manipulation}.
void nwbo_compute_manipulation_plan(const nwb_params* p, const double* boxes, int nb, const double* service13, const double* db,
                                    int n_db, const double* pose6, const double* start, int circular, const double* object,
                                    int execute, uint64_t seed, double* goal_approach, double* goal_manipulate,
                                    double* traj_a, double* tfs_a, double* traj_m, double* tfs_m, int cap, int32_t* stats) {
  PlanParams pp = make_params(p, boxes, nb);
  ManipulationPlanOut o = compute_manipulation_plan(pp, make_service(service13, p), db, n_db, pose6, start, circular != 0, object,
                                                    execute != 0, seed);
  if (o.goal_approach.num_found && goal_approach) std::memcpy(goal_approach, o.goal_approach.rows.data(), sizeof(double) * NWB_NJ);
  if (o.goal_manipulate.num_found && goal_manipulate) std::memcpy(goal_manipulate, o.goal_manipulate.rows.data(), sizeof(double) * NWB_NJ);
  int na = put_rows(o.approach.trajectory, traj_a, cap), nm = put_rows(o.manipulation.trajectory, traj_m, cap);
  if (tfs_a) std::memcpy(tfs_a, o.approach.time_from_start.data(), sizeof(double) * std::min(na, cap));
  if (tfs_m) std::memcpy(tfs_m, o.manipulation.time_from_start.data(), sizeof(double) * std::min(nm, cap));
  stats[0] = o.approach.success; stats[1] = o.manipulation.success; stats[2] = o.approach.num_nodes;
  stats[3] = o.manipulation.num_nodes; stats[4] = o.goal_approach.num_found; stats[5] = o.goal_manipulate.num_found;
  stats[6] = na; stats[7] = nm;
}

// generateTrajectory's time parameterisation (RP:1596-1618); limits taken from nwb_params.
void nwbo_time_parameterize(const nwb_params* p, const double* positions, int n_points, double* time_from_start) {
  TimeParamLimits L;
  for (int j = 0; j < NWB_NJ; ++j) {
    L.v_max[j] = p->joint_max_velocity[j];
    L.a_max[j] = p->joint_max_acceleration[j];
  }
  L.max_iterations = p->iptp_max_iterations;
  L.max_time_change = p->iptp_max_time_change;
  std::vector<double> path(positions, positions + (size_t)n_points * NWB_NJ);
  std::vector<double> st = time_parameterize(L, path);
  std::memcpy(time_from_start, st.data(), sizeof(double) * st.size());
}

}  // extern "C"
