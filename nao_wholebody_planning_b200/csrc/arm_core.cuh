// arm_core.cuh - per-thread (fp64) device functions for the right arm: hand position + direction
// inverse kinematics, the articulated-object (drawer / door) constraint and the ergonomics index.
//
// Reference functions replaced (file:line relative to the reference checkout):
//   ConstraintKinematics::computeRightArmIK            rrt_connect_planner/src/nao_constraint_kinematics.cpp:224-506
//   ConstraintKinematics::compute_ergonomics_index     nao_constraint_kinematics.cpp:612-659
//   ArticulatedObjectConstraints::right_hand_at_pose   rrt_connect_planner/src/articulated_object_constraints.cpp:204-262
//   ArticulatedObjectConstraints::compute_right_arm_IK articulated_object_constraints.cpp:313-449
//   RRT_Planner::enforce_MA_Constraints                rrt_connect_planner/src/rrt_planner.cpp:597-696
//   IKSolver::ik (generated, ikfast v55)               rrt_connect_planner/src/ikfast_TranslationDirection5D.cpp:341-1517
//
// The generated solver searches the roots of a degree-8 polynomial.  The arm it was generated for
// (its fk, ikfast_TranslationDirection5D.cpp:301-326) decouples, so this file solves the same problem
// in closed form, branch-free per solution and without iteration:
//   forearm axis x_f is perpendicular to the requested hand direction d  -> x_f = cos(phi) u + sin(phi) v
//   |elbow - shoulder| = |upper arm|                                     -> A cos(phi) + B sin(phi) = C   (2 roots)
//   shoulder pitch / roll from the elbow position                        (2 branches)
//   elbow yaw / roll from x_f in the upper-arm frame                     (2 branches)
//   wrist yaw from d
// i.e. the same (up to) 8 solutions.
#pragma once
#include "planner_core.cuh"

namespace nwb {

// the solver's arm model (torso frame)
constexpr double kArmShoulderY = -0.098, kArmShoulderZ = 0.100;
constexpr double kArmUpperX = 0.105, kArmUpperY = -0.015;
constexpr double kArmForearm = 0.12395, kArmHandOffset = 0.0159;
constexpr double kPi = 3.14159265358979323846;

// f.R <- f.R * RotZ(a)
__device__ __forceinline__ void drotz(DFrame& f, double a) {
  double s, c;
  sincos(a, &s, &c);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double u = f.r[3 * k + 0], v = f.r[3 * k + 1];
    f.r[3 * k + 0] = c * u + s * v;
    f.r[3 * k + 1] = c * v - s * u;
  }
}
__device__ __forceinline__ D3 tmul(const double* r, D3 v) {   // R^T v
  return d3(r[0] * v.x + r[3] * v.y + r[6] * v.z, r[1] * v.x + r[4] * v.y + r[7] * v.z, r[2] * v.x + r[5] * v.y + r[8] * v.z);
}
__device__ __forceinline__ double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 dscale(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double wrap_pi(double a) {
  if (a > kPi) a -= 2 * kPi;
  if (a <= -kPi) a += 2 * kPi;
  return a;
}

// Hand pose given in the right-sole frame -> the solver's torso frame (nao_constraint_kinematics.cpp:226-282,
// articulated_object_constraints.cpp:316-357): R = hip^T, pos = R (p - p_hip) - (0,0,0.085), dir = R d.
__device__ __forceinline__ void hand_target_in_torso(const DFrame& hip, const double* pose6, D3& pos, D3& dir) {
  double o[3], dd[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double r0 = hip.r[i], r1 = hip.r[3 + i], r2 = hip.r[6 + i];        // row i of hip^T
    double tmp = __dsub_rn(__dsub_rn(__dmul_rn(-r0, hip.p.x), __dmul_rn(r1, hip.p.y)), __dmul_rn(r2, hip.p.z));
    double tmp2 = __dadd_rn(__dadd_rn(__dmul_rn(r0, pose6[0]), __dmul_rn(r1, pose6[1])), __dmul_rn(r2, pose6[2]));
    o[i] = __dadd_rn(tmp, tmp2);
    dd[i] = __dadd_rn(__dadd_rn(__dmul_rn(r0, pose6[3]), __dmul_rn(r1, pose6[4])), __dmul_rn(r2, pose6[5]));
  }
  pos = d3(o[0], o[1], o[2] - 0.085);
  dir = d3(dd[0], dd[1], dd[2]);
}

// All joint vectors that put the grasp point at `pos` with hand axis `dir` (unit).  Calls
// visit(q[5]) for each solution; returns their number (0, 4 or 8; fewer at singular poses).
template <class Visit>
__device__ __forceinline__ int arm5d_ik(D3 pos, D3 d, Visit visit) {
  const double L = kArmForearm, h = kArmHandOffset;
  const double rho2 = kArmUpperX * kArmUpperX + kArmUpperY * kArmUpperY;
  const double rho = sqrt(rho2);
  const double delta = atan2(-kArmUpperY, kArmUpperX);
  D3 r = d3(pos.x, pos.y - kArmShoulderY, pos.z - kArmShoulderZ);
  D3 t = fabs(d.x) < 0.9 ? d3(1, 0, 0) : d3(0, 1, 0);
  D3 u = cross(t, d);
  u = dscale(u, 1.0 / sqrt(ddot(u, u)));
  D3 v = cross(d, u);
  const double ru = ddot(r, u), rv = ddot(r, v);
  const double A = -L * ru + h * rv, B = -L * rv - h * ru;
  const double C = 0.5 * (rho2 - ddot(r, r) - L * L - h * h);
  const double amp = sqrt(A * A + B * B);
  if (!(amp > 0) || fabs(C) > amp) return 0;
  const double phi0 = atan2(B, A), dphi = acos(C / amp);
  int n = 0;
#pragma unroll 1
  for (int s_phi = 0; s_phi < 2; ++s_phi) {
    double sp, cp;
    sincos(phi0 + (s_phi ? -dphi : dphi), &sp, &cp);
    D3 xf = dscale(u, cp) + dscale(v, sp);
    D3 e = r - dscale(xf, L) + dscale(cross(d, xf), h);
    const double sy = fmin(1.0, fmax(-1.0, e.y / rho));
    const double a1 = asin(sy);
#pragma unroll 1
    for (int s_sh = 0; s_sh < 2; ++s_sh) {
      const double q1 = delta + (s_sh ? kPi - a1 : a1);
      const double radial = rho * cos(q1 - delta);
      if (fabs(radial) < 1e-12) continue;
      const double sg = radial > 0 ? 1.0 : -1.0;
      const double q0 = atan2(-e.z * sg, e.x * sg);
      DFrame f01;
      dframe_identity(f01);
      droty(f01, q0);
      drotz(f01, q1);
      D3 f = tmul(f01.r, xf);
      const double c3 = fmin(1.0, fmax(-1.0, f.x));
      const double a3 = acos(c3);
#pragma unroll 1
      for (int s_el = 0; s_el < 2; ++s_el) {
        const double q3 = s_el ? -a3 : a3;
        const double s3 = sin(q3);
        if (fabs(s3) < 1e-12) continue;
        const double q2 = atan2(f.z / s3, f.y / s3);
        DFrame f03 = f01;
        drotx(f03, q2);
        drotz(f03, q3);
        D3 dl = tmul(f03.r, d);
        const double q4 = atan2(-dl.z, -dl.y);
        double q[5] = {wrap_pi(q0), wrap_pi(q1), wrap_pi(q2), wrap_pi(q3), wrap_pi(q4)};
        visit(q);
        ++n;
      }
    }
  }
  return n;
}

// chain_right_arm_ (hip frame -> grasp frame), nao_constraint_kinematics.cpp:39-47, appended to `f`;
// optionally records joint origins / axes for the Jacobian.  The final fixed segment
// (RPY(pi/2,0,0), (0.01,0,0)) is applied too, so f.r's third column is the hand z-axis.
__device__ __forceinline__ void right_arm_fk(const double* q5, DFrame& f, D3* origin, D3* axis) {
  dtrans(f, 0, 0, 0.182);
  dtrans(f, 0, -0.098, 0);
  if (origin) { origin[0] = f.p; axis[0] = col(f.r, 1); }
  droty(f, q5[0]);
  if (origin) { origin[1] = f.p; axis[1] = col(f.r, 2); }
  drotz(f, q5[1]); dtrans(f, 0.105, -0.015, 0);
  if (origin) { origin[2] = f.p; axis[2] = col(f.r, 0); }
  drotx(f, q5[2]);
  if (origin) { origin[3] = f.p; axis[3] = col(f.r, 2); }
  drotz(f, q5[3]); dtrans(f, 0.05595, 0, 0);
  if (origin) { origin[4] = f.p; axis[4] = col(f.r, 0); }
  drotx(f, q5[4]); dtrans(f, 0.05775, 0, -0.01231);
  dtrans(f, 0.01, 0, 0);
  drotx(f, 1.57079632679489661923);              // KDL::Rotation::RPY(M_PI_2, 0, 0)
}

// compute_ergonomics_index: RShoulderPitch * |"det"| with the shipped diagonal-product formula over
// [J J] (forward: 5 wrapped diagonals; "backward": one column multiplied down all six rows).
__device__ __forceinline__ double ergonomics_index(const double* q5) {
  DFrame f;
  dframe_identity(f);
  D3 o[5], a[5];
  right_arm_fk(q5, f, o, a);
  double J[6][5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    D3 v = cross(a[k], f.p - o[k]);
    J[0][k] = v.x; J[1][k] = v.y; J[2][k] = v.z;
    J[3][k] = a[k].x; J[4][k] = a[k].y; J[5][k] = a[k].z;
  }
  double sum_fw = 0.0, sum_bw = 0.0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    double pf = J[0][i % 5];
#pragma unroll
    for (int r = 1; r < 6; ++r) pf = __dmul_rn(pf, J[r][(i + r) % 5]);
    sum_fw = __dadd_rn(sum_fw, pf);
    const int c = 4 - i;                                    // column new_cols_number - (1 + i) of [J J]
    double pb = J[0][c];
#pragma unroll
    for (int r = 1; r < 6; ++r) pb = __dmul_rn(pb, J[r][c]);
    sum_bw = __dsub_rn(sum_bw, pb);
  }
  return __dmul_rn(q5[0], fabs(__dsub_rn(sum_fw, sum_bw)));
}

// ChainIkSolverPos_NR::CartToJnt on the right-arm chain (base = hip frame), as the position-only branch of
// computeRightArmIK constructs it (nao_constraint_kinematics.cpp:311-317).  Same iteration as left_leg_ik_nr.
__device__ __forceinline__ bool right_arm_ik_nr(const DFrame& target, double* q /*[5] in: seed*/, int maxiter, double eps,
                                                double pinv_eps) {
  double V[5][5];                                         // warm start of the Jacobi SVD across iterations
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int b = 0; b < 5; ++b) V[a][b] = (a == b) ? 1.0 : 0.0;
  int i;
  for (i = 0; i < maxiter; ++i) {
    DFrame f;
    dframe_identity(f);
    D3 o[5], a[5];
    right_arm_fk(q, f, o, a);
    double tw[6];
    frame_diff(f, target, tw);
    double W[6][5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      D3 v = cross(a[k], f.p - o[k]);
      W[0][k] = v.x; W[1][k] = v.y; W[2][k] = v.z;
      W[3][k] = a[k].x; W[4][k] = a[k].y; W[5][k] = a[k].z;
    }
    double dq[5];
    pinv_solve(W, tw, pinv_eps, dq, V);
#pragma unroll
    for (int k = 0; k < 5; ++k) q[k] += dq[k];
    bool conv = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) conv = conv && (eps > tw[k]) && (tw[k] > -eps);
    if (conv) break;
  }
  return i != maxiter;
}

constexpr uint32_t kStreamArmSeed = 3;

// computeRightArmIK (nao_constraint_kinematics.cpp:224-506).  hip = right-leg FK of the sample.
// Direction != 0: the generated solver's solutions, strict limits, largest ergonomics index (> 0.0).
// Direction == (0,0,0) exactly (nao_constraint_kinematics.cpp:291-372): KDL NR towards Frame(position in the hip
// frame) - identity rotation - from up to 5 random seeds inside the arm limits (stream kStreamArmSeed of the
// sample, slot 5 * attempt + joint); the ergonomics index stays untouched (0).
__device__ __forceinline__ bool goal_right_arm_ik(const DFrame& hip, const double* pose6, const PlanConst& pc, uint64_t seed,
                                                  uint64_t sample_index, double* out5, double* ergonomics) {
  const double* mn5 = pc.jmin + 10;
  const double* mx5 = pc.jmax + 10;
  *ergonomics = 0.0;
  if (pose6[3] == 0.0 && pose6[4] == 0.0 && pose6[5] == 0.0) {
    DFrame target;
    dframe_identity(target);
    double o[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double r0 = hip.r[i], r1 = hip.r[3 + i], r2 = hip.r[6 + i];
      double tmp = __dsub_rn(__dsub_rn(__dmul_rn(-r0, hip.p.x), __dmul_rn(r1, hip.p.y)), __dmul_rn(r2, hip.p.z));
      double tmp2 = __dadd_rn(__dadd_rn(__dmul_rn(r0, pose6[0]), __dmul_rn(r1, pose6[1])), __dmul_rn(r2, pose6[2]));
      o[i] = __dadd_rn(tmp, tmp2);
    }
    target.p = d3(o[0], o[1], o[2]);
    bool ok = false;
#pragma unroll 1
    for (int attempt = 0; attempt < 5 && !ok; ++attempt) {
      double q[5];
#pragma unroll 1
      for (int blk = 0; blk < 2; ++blk) {                  // slots 5 * attempt .. 5 * attempt + 4 span at most two counter blocks
        const int first_slot = 5 * attempt, b = first_slot / 4 + blk;
        uint32_t r[4];
        philox4x32_10((uint32_t)sample_index, (uint32_t)(sample_index >> 32), (uint32_t)b, kStreamArmSeed, (uint32_t)seed,
                      (uint32_t)(seed >> 32), r);
        for (int l = 0; l < 4; ++l) {
          const int j = 4 * b + l - first_slot;
          if (j >= 0 && j < 5) q[j] = __dadd_rn(__dmul_rn(mx5[j] - mn5[j], u01(r[l])), mn5[j]);
        }
      }
      // slots 5a..5a+4 may need a third block only if they straddle two boundaries: 5 consecutive slots never do
      if (right_arm_ik_nr(target, q, pc.ik_maxiter, pc.ik_eps, pc.pinv_eps) && limits_ok(q, mn5, mx5, 5)) {
        ok = true;
#pragma unroll
        for (int j = 0; j < 5; ++j) out5[j] = q[j];
      }
    }
    return ok;
  }
  D3 pos, dir;
  hand_target_in_torso(hip, pose6, pos, dir);
  double best = 0.0;
  bool found = false;
  arm5d_ik(pos, dir, [&](const double* q) {
    if (!limits_ok(q, mn5, mx5, 5)) return;
    const double e = ergonomics_index(q);
    if (e > best) {
      best = e;
      found = true;
#pragma unroll
      for (int j = 0; j < 5; ++j) out5[j] = q[j];
    }
  });
  *ergonomics = best;
  return found;
}

// The hand trajectory of the articulated object: n_pts rows of (position, hand z-axis).
struct MaView {
  const double* pts;      // device pointer, [n_pts][6]; nullptr = no constraint
  int n_pts;
};

// right_hand_at_pose: FK of the sole->hand chain; 5 mm per position component, 0.007 per axis component.
__device__ __forceinline__ bool right_hand_at_pose(const double* target6, const DFrame& hip, const double* arm5) {
  DFrame f = hip;
  right_arm_fk(arm5, f, nullptr, nullptr);
  const double e0 = fabs(f.p.x - target6[0]), e1 = fabs(f.p.y - target6[1]), e2 = fabs(f.p.z - target6[2]);
  const double e3 = fabs(f.r[2] - target6[3]), e4 = fabs(f.r[5] - target6[4]), e5 = fabs(f.r[8] - target6[5]);
  double max_pos = 0.0, max_dir = 0.0;
  if (e0 > max_pos) max_pos = e0;
  if (e1 > max_pos) max_pos = e1;
  if (e2 > max_pos) max_pos = e2;
  if (e3 > max_dir) max_dir = e3;
  if (e4 > max_dir) max_dir = e4;
  if (e5 > max_dir) max_dir = e5;
  if (max_pos > 0.005) return false;
  if (max_dir > 0.007) return false;
  return true;
}

// enforce_MA_Constraints: q (after the double-support projection) gets its right arm re-solved when
// the hand is not at the desired trajectory point; the solution closest (L1) to q_near's arm wins.
__device__ __forceinline__ int enforce_ma_constraints(const MaView& ma, bool start_tree, int near_hand_index, const double* q_near,
                                                      double* q, const PlanConst& pc) {
  int desired;
  if (start_tree && near_hand_index <= ma.n_pts - 2) desired = near_hand_index + 1;
  else if (!start_tree && near_hand_index > 0) desired = near_hand_index - 1;
  else desired = near_hand_index;
  double target[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) target[j] = ma.pts[(size_t)desired * 6 + j];
  double rl[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) rl[i] = -q[i];
  DFrame hip;
  right_leg_fk(rl, hip);
  if (right_hand_at_pose(target, hip, q + 10)) return NWB_REACHED;
  D3 pos, dir;
  hand_target_in_torso(hip, target, pos, dir);
  double best = 10000.0, sol[5];
  arm5d_ik(pos, dir, [&](const double* s) {
    if (!limits_ok(s, pc.jmin + 10, pc.jmax + 10, 5)) return;
    double dist = 0.0;
#pragma unroll
    for (int j = 0; j < 5; ++j) dist = __dadd_rn(dist, fabs(q_near[10 + j] - s[j]));
    if (dist < best) {
      best = dist;
#pragma unroll
      for (int j = 0; j < 5; ++j) sol[j] = s[j];
    }
  });
  if (!(best < 1000)) return NWB_TRAPPED;
#pragma unroll
  for (int j = 0; j < 5; ++j) q[10 + j] = sol[j];
  return NWB_ADVANCED;
}

// cart_hand_pose_index of a node appended behind q_near (addConfigtoTree, rrt_planner.cpp:757-767;
// note `< n_pts - 2` here against `<= n_pts - 2` in enforce_MA_Constraints: kept as shipped)
__device__ __forceinline__ int next_hand_index(const MaView& ma, bool start_tree, int near_hand_index) {
  if (start_tree && near_hand_index < ma.n_pts - 2) return near_hand_index + 1;
  if (!start_tree && near_hand_index > 0) return near_hand_index - 1;
  return near_hand_index;
}

}  // namespace nwb
