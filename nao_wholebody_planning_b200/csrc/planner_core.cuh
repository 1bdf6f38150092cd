// planner_core.cuh - per-thread (fp64) device functions of the local planner: steering, joint limits,
// double-support check and the left-leg Newton-Raphson / pseudo-inverse IK.
//
// Reference functions replaced (file:line relative to the reference checkout):
//   RRT_Planner::generate_q_new            rrt_connect_planner/src/rrt_planner.cpp:424-494
//   RRT_Planner::enforce_DS_Constraints    rrt_connect_planner/src/rrt_planner.cpp:498-593
//   LocalPlanner::checkJointLimits         rrt_connect_planner/src/local_planner.cpp:305-371
//   LocalPlanner::checkLeftLegPose         rrt_connect_planner/src/local_planner.cpp:244-299
//   LocalPlanner::getHipPose / getDesiredLeftLegPose / computeLeftLegIK   local_planner.cpp:59-236
//   ConstraintKinematics::getHipPose / computeLeftLegIK / joint_limits_respected
//                                          rrt_connect_planner/src/nao_constraint_kinematics.cpp:71-197,512-608
// and the orocos-KDL solvers those construct (ChainFkSolverPos_recursive, ChainJntToJacSolver,
// ChainIkSolverVel_pinv eps 1e-5, ChainIkSolverPos_NR maxiter 100 eps 1e-3).
//
// Double precision on purpose: B200 keeps a full-rate FP64 pipe, the NR iteration count depends on
// comparisons against 1e-3 / 1e-5 thresholds, and running it in fp64 makes the device take the same
// branch as the CPU path except within ~1e-12 of a threshold.  The leg chains are compiled in
// (local_planner.cpp:18-47): RotX/RotY joints with z- or y- offsets, so FK and the Jacobian are
// closed-form products of planar rotations instead of generic segment walks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nwb.h"
#include "nao_model_gen.h"

namespace nwb {

struct PlanConst {            // passed by value to the planner kernels
  double jmin[NWB_NJ], jmax[NWB_NJ];
  double weights[NWB_NJ];
  double step;
  double ds_tolerance, ik_eps, pinv_eps;
  int ik_maxiter;
  unsigned quirks;
};

struct D3 {
  double x, y, z;
};
struct DFrame {               // rotation row-major + translation
  double r[9];
  D3 p;
};

__device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 cross(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ D3 rmul(const double* r, D3 v) {
  return d3(r[0] * v.x + r[1] * v.y + r[2] * v.z, r[3] * v.x + r[4] * v.y + r[5] * v.z, r[6] * v.x + r[7] * v.y + r[8] * v.z);
}
__device__ __forceinline__ D3 col(const double* r, int c) { return d3(r[c], r[3 + c], r[6 + c]); }

__device__ __forceinline__ void dframe_identity(DFrame& f) {
  f.r[0] = 1; f.r[1] = 0; f.r[2] = 0; f.r[3] = 0; f.r[4] = 1; f.r[5] = 0; f.r[6] = 0; f.r[7] = 0; f.r[8] = 1;
  f.p = d3(0, 0, 0);
}
// f.R <- f.R * RotX(a) / RotY(a)  (KDL::Rotation::RotX/RotY); the _sc forms take sin / cos of the angle
__device__ __forceinline__ void drotx_sc(DFrame& f, double s, double c) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double u = f.r[3 * k + 1], v = f.r[3 * k + 2];
    f.r[3 * k + 1] = c * u + s * v;
    f.r[3 * k + 2] = c * v - s * u;
  }
}
__device__ __forceinline__ void droty_sc(DFrame& f, double s, double c) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double u = f.r[3 * k + 0], v = f.r[3 * k + 2];
    f.r[3 * k + 0] = c * u - s * v;
    f.r[3 * k + 2] = s * u + c * v;
  }
}
__device__ __forceinline__ void drotx(DFrame& f, double a) {
  double s, c;
  sincos(a, &s, &c);
  drotx_sc(f, s, c);
}
__device__ __forceinline__ void droty(DFrame& f, double a) {
  double s, c;
  sincos(a, &s, &c);
  droty_sc(f, s, c);
}
__device__ __forceinline__ void dtrans(DFrame& f, double x, double y, double z) { f.p = f.p + rmul(f.r, d3(x, y, z)); }

// chain_right_leg_ (sole -> hip), local_planner.cpp:18-23.  legs5 = (-RAnkleRoll, -RAnklePitch, -RKneePitch, -RHipPitch, -RHipRoll)
__device__ __forceinline__ void right_leg_fk(const double* legs5, DFrame& f) {
  dframe_identity(f);
  f.p = d3(0, 0, 0.04519);
  drotx(f, legs5[0]);
  droty(f, legs5[1]); dtrans(f, 0, 0, 0.1029);
  droty(f, legs5[2]); dtrans(f, 0, 0, 0.1);
  droty(f, legs5[3]);
  drotx(f, legs5[4]); dtrans(f, 0, 0.05, 0);
}
// the same two chains with the sines / cosines of their five joint angles supplied (computed on five lanes of a warp)
__device__ __forceinline__ void right_leg_fk_sc(const double* sn, const double* cs, DFrame& f) {
  dframe_identity(f);
  f.p = d3(0, 0, 0.04519);
  drotx_sc(f, sn[0], cs[0]);
  droty_sc(f, sn[1], cs[1]); dtrans(f, 0, 0, 0.1029);
  droty_sc(f, sn[2], cs[2]); dtrans(f, 0, 0, 0.1);
  droty_sc(f, sn[3], cs[3]);
  drotx_sc(f, sn[4], cs[4]); dtrans(f, 0, 0.05, 0);
}
__device__ __forceinline__ void left_leg_fk_sc(const double* sn, const double* cs, DFrame& f) {
  dtrans(f, 0, 0.05, 0);
  drotx_sc(f, sn[0], cs[0]);
  droty_sc(f, sn[1], cs[1]); dtrans(f, 0, 0, -0.1);
  droty_sc(f, sn[2], cs[2]); dtrans(f, 0, 0, -0.1029);
  droty_sc(f, sn[3], cs[3]);
  drotx_sc(f, sn[4], cs[4]); dtrans(f, 0, 0, -0.04519);
}

// chain_left_leg_ (hip -> sole), local_planner.cpp:26-31, appended to `f`; optionally records the
// joint origins / axes (base frame of `f`'s parent) for the Jacobian.
__device__ __forceinline__ void left_leg_fk(const double* ll5, DFrame& f, D3* origin, D3* axis) {
  dtrans(f, 0, 0.05, 0);
  if (origin) { origin[0] = f.p; axis[0] = col(f.r, 0); }
  drotx(f, ll5[0]);
  if (origin) { origin[1] = f.p; axis[1] = col(f.r, 1); }
  droty(f, ll5[1]); dtrans(f, 0, 0, -0.1);
  if (origin) { origin[2] = f.p; axis[2] = col(f.r, 1); }
  droty(f, ll5[2]); dtrans(f, 0, 0, -0.1029);
  if (origin) { origin[3] = f.p; axis[3] = col(f.r, 1); }
  droty(f, ll5[3]);
  if (origin) { origin[4] = f.p; axis[4] = col(f.r, 0); }
  drotx(f, ll5[4]); dtrans(f, 0, 0, -0.04519);
}

// KDL::Rotation::GetRot of R = A^T B (rotation vector axis*angle; GetRotAngle with epsilon 1e-6)
__device__ __forceinline__ D3 rot_vector(const double* R) {
  const double eps = 1e-6;
  double ca = (R[0] + R[4] + R[8] - 1) / 2.0;
  double t = eps * eps / 2.0;
  if (ca > 1 - t) return d3(0, 0, 0);
  if (ca < -1 + t) {
    double x = sqrt((R[0] + 1.0) / 2), y = sqrt((R[4] + 1.0) / 2), z = sqrt((R[8] + 1.0) / 2);
    if (R[2] < 0) x = -x;
    if (R[7] < 0) y = -y;
    if (x * y * R[1] < 0) x = -x;
    const double pi = 3.14159265358979323846;
    return d3(x * pi, y * pi, z * pi);
  }
  double ax = R[7] - R[5], ay = R[2] - R[6], az = R[3] - R[1];
  double mod = sqrt(ax * ax + ay * ay + az * az);
  double angle = atan2(mod / 2, ca);
  return d3(ax / mod * angle, ay / mod * angle, az / mod * angle);
}

// KDL::diff(F_a, F_b) = (p_b - p_a, R_a * rotvec(R_a^T R_b))
__device__ __forceinline__ void frame_diff(const DFrame& a, const DFrame& b, double* tw) {
  double rel[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) rel[3 * i + j] = a.r[i] * b.r[j] + a.r[3 + i] * b.r[3 + j] + a.r[6 + i] * b.r[6 + j];
  D3 w = rmul(a.r, rot_vector(rel));
  tw[0] = b.p.x - a.p.x; tw[1] = b.p.y - a.p.y; tw[2] = b.p.z - a.p.z;
  tw[3] = w.x; tw[4] = w.y; tw[5] = w.z;
}

// One Hestenes rotation of columns (P, Q) of W (6x5) and V (5x5).  All indices are compile-time constants.
// t = tan(theta) of the rotation that orthogonalises the two columns, written without the intermediate
// zeta = (beta - alpha) / (2 gamma):  t = sgn(zeta) 2|gamma| / (|d| + sqrt(d^2 + 4 gamma^2)),  d = beta - alpha
// (one square root, one division, one reciprocal square root per rotation).
template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(double (&W)[6][5], double (&V)[5][5], bool& rotated) {
  double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    alpha += W[r][P] * W[r][P];
    beta += W[r][Q] * W[r][Q];
    gamma += W[r][P] * W[r][Q];
  }
  // converged pair: |gamma| <= 1e-15 sqrt(alpha beta)   (compared squared: no root, no division)
  if (gamma * gamma > 1e-30 * (alpha * beta)) {
    rotated = true;
    const double d = beta - alpha;
    const double sg = ((d >= 0) == (gamma >= 0) || d == 0) ? 1.0 : -1.0;
    const double t = sg * (2.0 * fabs(gamma)) / (fabs(d) + sqrt(d * d + 4.0 * gamma * gamma));
    const double c = rsqrt(1.0 + t * t), s = c * t;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double wp = W[r][P], wq = W[r][Q];
      W[r][P] = c * wp - s * wq;
      W[r][Q] = s * wp + c * wq;
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const double vp = V[r][P], vq = V[r][Q];
      V[r][P] = c * vp - s * vq;
      V[r][Q] = s * vp + c * vq;
    }
  }
}

// qdot = pinv(J) * twist with singular values below eps zeroed (ChainIkSolverVel_pinv).
// J is 6x5 (column i = (a_i x (p_tip - o_i), a_i)).  One-sided Jacobi SVD, W = J V = U S.  A sweep visits the
// 10 column pairs in a round-robin order - five rounds of two DISJOINT pairs - so the two rotations of a
// round are independent instruction streams: this code runs one warp per scheduler with nothing else to
// hide the fp64 dependency chains behind.
// WARM START: V_io carries the right singular vectors from one Newton-Raphson iteration to the next.  The
// Jacobian moves little between iterations, so J V_prev is already nearly orthogonal and one or two sweeps
// finish the job instead of six; the first call of a solve passes the identity.  The decomposition reached
// is the same (the pseudo-inverse is unique), only the route is shorter.
static __device__ __noinline__ void pinv_solve(const double (*W_in)[5], const double* tw_in, double eps, double* qdot,
                                               double (*V_io)[5]) {
  // W_in / tw_in / V_io arrive through memory (the function is not inlined): pull them into registers once
  double W[6][5], tw[6], V[5][5];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) V[i][j] = V_io[i][j];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    tw[r] = tw_in[r];
#pragma unroll
    for (int c = 0; c < 5; ++c) {                        // W = J V
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < 5; ++m) acc += W_in[r][m] * V[m][c];
      W[r][c] = acc;
    }
  }
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    jacobi_rotate<1, 4>(W, V, rotated); jacobi_rotate<2, 3>(W, V, rotated);
    jacobi_rotate<0, 2>(W, V, rotated); jacobi_rotate<3, 4>(W, V, rotated);
    jacobi_rotate<0, 4>(W, V, rotated); jacobi_rotate<1, 3>(W, V, rotated);
    jacobi_rotate<0, 1>(W, V, rotated); jacobi_rotate<2, 4>(W, V, rotated);
    jacobi_rotate<0, 3>(W, V, rotated); jacobi_rotate<1, 2>(W, V, rotated);
    if (!rotated) break;
  }
  // tmp_i = (u_i . tw) / s_i  with u_i = W[:,i]/s_i  ->  (W[:,i] . tw) / s_i^2 ; zero if s_i < eps
  double tmp[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    double n2 = 0, dotv = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      n2 += W[r][i] * W[r][i];
      dotv += W[r][i] * tw[r];
    }
    double sv = sqrt(n2);
    tmp[i] = (sv < eps) ? 0.0 : (dotv / sv) * (1.0 / sv);
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    double s = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      s += V[i][j] * tmp[j];
      V_io[i][j] = V[i][j];
    }
    qdot[i] = s;
  }
}

// Exact rejection of left-sole targets the Newton-Raphson solve can never converge to.  Convergence at iteration i
// means that the frame of q_i differs from the target by a twist with EVERY component inside (-eps, eps)
// (KDL::Equal(delta_twist, Zero, eps)): a translation dp with |dp_k| < eps and a base-frame rotation vector w with
// |w_k| < eps; both have norm <= sqrt(3) eps =: t.  With c = (0, 0.05, 0) the hip joint centre, the chain is
//   sole = c + Rx(q0) [ Ry(q1)(0,0,-0.1) + Ry(q1+q2)(0,0,-0.1029) ] + R (0,0,-0.04519),   R = Rx(q0) Ry(q1+q2+q3) Rx(q4),
// so for EVERY joint vector (5 joints, 6 pose coordinates: one equality and two inequalities are left over)
//   (1) |sole - c| <= L = 0.24809,
//   (2) the ankle A = sole + 0.04519 R e_z satisfies |A - c| <= 0.2029,
//   (3) F := u . (A - c) = 0 with u = e_x x R e_x: the leg plane's normal n = Rx(q0) e_y is normal to e_x and to R e_x,
//       so n is parallel to u (or u = 0), and A - c lies in the leg plane.
// A target that CAN be converged to is within the tolerance box of such a frame.  (1) and (2) with the norms carried
// through: |sole_t - c| <= L + t, |A_t - c| <= 0.2029 + (1 + 0.04519) t.  (3) to first order over the box, with a
// rigorous remainder: perturbing the target by (dp, w) changes F by
//   -u . dp - w . g,   g = 0.04519 (z x u) + x x ((A - c) x e_x),   x = R e_x, z = R e_z,
// whose largest magnitude over the box is eps (|u|_1 + |g|_1); what the expansion leaves out is bounded by
//   R2 = t^2 (1 + f) + h (|A - c| + t (1 + f) + f h) + (|u| + t) f h,   f = 0.04519, h = 0.51 t^2
// (a rotation by |w| <= t moves a unit vector by w x v up to |w|^2/2 + |w|^3/6 <= h).  F must vanish somewhere in the
// box, so |F(target)| <= eps (|u|_1 + |g|_1) + R2 for every target the solve can converge to.
// Anything else runs all `maxiter` iterations and fails.  On uniform samples (oracle, 3 x 10^5): 37 % rejected by (1),
// 7 % by (2), 39 % by (3); 96 % of all failing solves; none of the 37 842 converging solves is rejected, they reach at
// most 0.925 of the bound of (3) - the L1 bound is attained only at the corners of the box.  (Round 1 bounded (3) with
// the Euclidean norms throughout, a bound 1.6 x wider: 87 % of the failing solves.)  1e-4 relative slack and 1e-6 on
// the lengths cover rounding.  The CPU oracle runs the full iteration for every target - the parity tests compare the
// verdicts, and tests/test_leg_rejection_rule_cpu.py restates the rule in numpy against the oracle on every CPU run.
__device__ __forceinline__ bool left_leg_target_feasible(const DFrame& target, double eps) {
  const double t = 1.7320508075688772 * eps, foot = 0.04519, legs = 0.1 + 0.1029;
  const double dx = target.p.x, dy = target.p.y - 0.05, dz = target.p.z;
  const double reach = legs + foot + t + 1e-6;
  if (dx * dx + dy * dy + dz * dz > reach * reach) return false;
  const double ax = dx + foot * target.r[2], ay = dy + foot * target.r[5], az = dz + foot * target.r[8];
  const double da = (1.0 + foot) * t, ankle = legs + da + 1e-6;
  const double an2 = ax * ax + ay * ay + az * az;
  if (an2 > ankle * ankle) return false;
  const double xx = target.r[0], xy = target.r[3], xz = target.r[6];   // x = R e_x
  const double zx = target.r[2], zy = target.r[5], zz = target.r[8];   // z = R e_z
  const double uy = -xz, uz = xy;                                       // u = e_x x x = (0, -x_z, x_y)
  const double F0 = uy * ay + uz * az;
  const double gx = foot * (zy * uz - zz * uy) - (xy * ay + xz * az);
  const double gy = xx * ay - foot * (zx * uz);
  const double gz = xx * az + foot * (zx * uy);
  const double B1 = eps * (fabs(uy) + fabs(uz) + fabs(gx) + fabs(gy) + fabs(gz));
  const double h = 0.51 * t * t;
  const double R2 = t * t * (1.0 + foot) + h * (sqrt(an2) + da + foot * h) + (sqrt(uy * uy + uz * uz) + t) * foot * h;
  return !(fabs(F0) > 1.0001 * (B1 + R2) + 1e-12);
}

// Iterations [i0, i1) of ChainIkSolverPos_NR::CartToJnt on the left-leg chain (base = hip frame).  The update is
// applied BEFORE the convergence test.  Returns the index of the iteration whose test succeeded, or -1; q and V (the
// right singular vectors carried from one iteration to the next) hold the state to resume from.
__device__ __forceinline__ int left_leg_ik_nr_span(const DFrame& target, double* q, double (*V)[5], int i0, int i1, double eps,
                                                   double pinv_eps) {
  for (int i = i0; i < i1; ++i) {
    DFrame f;
    dframe_identity(f);
    D3 o[5], a[5];
    left_leg_fk(q, f, o, a);
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
    if (conv) return i;
  }
  return -1;
}
__device__ __forceinline__ void nr_identity(double (*V)[5]) {
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int b = 0; b < 5; ++b) V[a][b] = (a == b) ? 1.0 : 0.0;
}

// The whole solve.  Returns true on convergence; q holds the result.
__device__ __forceinline__ bool left_leg_ik_nr(const DFrame& target, double* q /*[5] in: seed*/, int maxiter, double eps,
                                               double pinv_eps, int* iters_out) {
  // exact early out (see left_leg_target_feasible): the verdict and the iteration count of the full loop
  if (!left_leg_target_feasible(target, eps)) {
    if (iters_out) *iters_out = maxiter;
    return false;
  }
  double V[5][5];
  nr_identity(V);
  const int i = left_leg_ik_nr_span(target, q, V, 0, maxiter, eps, pinv_eps);
  if (iters_out) *iters_out = (i >= 0) ? i + 1 : maxiter;
  return i >= 0;
}

// target = (R^T, R^T ((0,0.1,0) - p_hip))   (getDesiredLeftLegPose, local_planner.cpp:59-110)
__device__ __forceinline__ void desired_left_leg_pose(const DFrame& hip, DFrame& t) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) t.r[3 * i + j] = hip.r[3 * j + i];
  const double fx = nwbm::kLeftSoleTarget[0], fy = nwbm::kLeftSoleTarget[1], fz = nwbm::kLeftSoleTarget[2];
  double o[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double tmp = -t.r[3 * i] * hip.p.x - t.r[3 * i + 1] * hip.p.y - t.r[3 * i + 2] * hip.p.z;
    double tmp2 = t.r[3 * i] * fx + t.r[3 * i + 1] * fy + t.r[3 * i + 2] * fz;
    o[i] = tmp + tmp2;
  }
  t.p = d3(o[0], o[1], o[2]);
}

// strict joint-limit test (checkJointLimits / joint_limits_respected); swapped bounds tolerated
__device__ __forceinline__ bool limits_ok(const double* q, const double* mn, const double* mx, int n) {
  bool ok = true;
  for (int i = 0; i < n; ++i) {
    double lo = fmin(mn[i], mx[i]), hi = fmax(mn[i], mx[i]);
    if (mn[i] != mx[i]) ok = ok && (q[i] > lo) && (q[i] < hi);
  }
  return ok;
}

// RRT_Planner::generate_q_new.  Returns NWB_ADVANCED or NWB_REACHED.
__device__ __forceinline__ int generate_q_new(const double* q_near, const double* q_rand, const PlanConst& pc, double* q_new) {
  double d[NWB_NJ], sum = 0;
  for (int i = 0; i < NWB_NJ; ++i) {
    d[i] = q_rand[i] - q_near[i];
    sum = __dadd_rn(sum, __dmul_rn(d[i], d[i]));
  }
  double norm = sqrt(sum);
  double inc[NWB_NJ], sum_inc = 0;
  for (int j = 0; j < NWB_NJ; ++j) {
    double dir = d[j] / norm;
    inc[j] = __dmul_rn(pc.step, __dmul_rn(dir, pc.weights[j]));
    sum_inc = __dadd_rn(sum_inc, __dmul_rn(inc[j], inc[j]));
  }
  double norm_inc = sqrt(sum_inc);
  if (norm_inc < norm) {
    for (int i = 0; i < NWB_NJ; ++i) q_new[i] = q_near[i] + inc[i];
    return NWB_ADVANCED;
  }
  for (int i = 0; i < NWB_NJ; ++i) q_new[i] = q_rand[i];
  return NWB_REACHED;
}

// RRT_Planner::enforce_DS_Constraints.  q is modified in place (left leg) when the IK is needed.
// Returns NWB_REACHED (already in double support), NWB_ADVANCED (left leg re-solved), NWB_TRAPPED.
__device__ __forceinline__ int enforce_ds_constraints(double* q, const PlanConst& pc, int* iters_out) {
  if (iters_out) *iters_out = 0;
  if (!limits_ok(q, pc.jmin, pc.jmax, NWB_NJ)) return NWB_TRAPPED;
  double rl[5], ll[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    rl[i] = -q[i];
    ll[i] = q[16 + i];
  }
  DFrame hip;
  right_leg_fk(rl, hip);
  // checkLeftLegPose: FK of the 10-joint legs chain, signed error (quirk C2) or absolute error
  DFrame foot = hip;
  left_leg_fk(ll, foot, nullptr, nullptr);
  const double des[3] = {nwbm::kLeftSoleTarget[0], nwbm::kLeftSoleTarget[1], nwbm::kLeftSoleTarget[2]};
  const double act[3] = {foot.p.x, foot.p.y, foot.p.z};
  double max_err = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double e = des[i] - act[i];
    if (!(pc.quirks & NWB_QUIRK_SIGNED_FOOT_ERROR)) e = fabs(e);
    if (e > max_err) max_err = e;
  }
  if (!(max_err > pc.ds_tolerance)) return NWB_REACHED;
  DFrame target;
  desired_left_leg_pose(hip, target);
  if (!left_leg_ik_nr(target, ll, pc.ik_maxiter, pc.ik_eps, pc.pinv_eps, iters_out)) return NWB_TRAPPED;
  if (!limits_ok(ll, pc.jmin + 16, pc.jmax + 16, 5)) return NWB_TRAPPED;
#pragma unroll
  for (int i = 0; i < 5; ++i) q[16 + i] = ll[i];
  return NWB_ADVANCED;
}

// ---- generate_q_new + enforce_DS_Constraints by one WARP for one pair (the RRT kernels) --------------------------
// The arithmetic is that of the two functions above, operation for operation (the parity tests compare iteration counts
// of whole searches); what changes is who does it.  Every lane forms the two serial sums itself (same order, same
// roundings), lane j takes the division and the joint-limit test of joint j, lanes 0-9 the fp64 sincos of the ten leg
// angles (~100 instructions each, the largest serial item of the walk), and lane 0 keeps the matrix chain, the foot
// error and - where the foot is off target - the Newton-Raphson solve.
struct WarpSteer {
  double near[NWB_NJ];   // q_near (in)
  double inc[NWB_NJ];    // scratch
  double sn[10], cs[10]; // sines / cosines of (-q[0..4], q[16..20])
};
// tgt = q_rand / q_connect, q = result (both 21 doubles in shared memory).  Returns NWB_TRAPPED / ADVANCED / REACHED, uniform.
__device__ __forceinline__ int warp_steer_project(const PlanConst& pc, WarpSteer& ws, const double* tgt, double* q, int lane) {
  const unsigned FULL = 0xffffffffu;
  double sum = 0;
  for (int i = 0; i < NWB_NJ; ++i) {
    const double di = tgt[i] - ws.near[i];
    sum = __dadd_rn(sum, __dmul_rn(di, di));
  }
  const double norm = sqrt(sum);
  if (lane < NWB_NJ) {
    const double dir = (tgt[lane] - ws.near[lane]) / norm;
    ws.inc[lane] = __dmul_rn(pc.step, __dmul_rn(dir, pc.weights[lane]));
  }
  __syncwarp();
  double sum_inc = 0;
  for (int j = 0; j < NWB_NJ; ++j) sum_inc = __dadd_rn(sum_inc, __dmul_rn(ws.inc[j], ws.inc[j]));
  const double norm_inc = sqrt(sum_inc);
  int st = norm_inc < norm ? NWB_ADVANCED : NWB_REACHED;
  bool ok = true;
  if (lane < NWB_NJ) {
    const double v = st == NWB_ADVANCED ? ws.near[lane] + ws.inc[lane] : tgt[lane];
    q[lane] = v;
    const double mn = pc.jmin[lane], mx = pc.jmax[lane];
    const double lo = fmin(mn, mx), hi = fmax(mn, mx);
    if (mn != mx) ok = (v > lo) && (v < hi);
  }
  if (!__all_sync(FULL, ok)) return NWB_TRAPPED;
  if (lane < 5 || (lane >= 16 && lane < NWB_NJ)) {        // right-leg chain runs on -q[0..4], left-leg chain on q[16..20]
    const int slot = lane < 5 ? lane : 5 + lane - 16;
    const double a = lane < 5 ? -q[lane] : q[lane];
    sincos(a, &ws.sn[slot], &ws.cs[slot]);
  }
  __syncwarp();
  int ds = NWB_REACHED;
  if (lane == 0) {
    DFrame hip;
    right_leg_fk_sc(ws.sn, ws.cs, hip);
    DFrame foot = hip;
    left_leg_fk_sc(ws.sn + 5, ws.cs + 5, foot);
    const double des[3] = {nwbm::kLeftSoleTarget[0], nwbm::kLeftSoleTarget[1], nwbm::kLeftSoleTarget[2]};
    const double act[3] = {foot.p.x, foot.p.y, foot.p.z};
    double max_err = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double e = des[i] - act[i];
      if (!(pc.quirks & NWB_QUIRK_SIGNED_FOOT_ERROR)) e = fabs(e);
      if (e > max_err) max_err = e;
    }
    if (max_err > pc.ds_tolerance) {
      double ll[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) ll[i] = q[16 + i];
      DFrame target;
      desired_left_leg_pose(hip, target);
      if (!left_leg_ik_nr(target, ll, pc.ik_maxiter, pc.ik_eps, pc.pinv_eps, nullptr) || !limits_ok(ll, pc.jmin + 16, pc.jmax + 16, 5)) {
        ds = NWB_TRAPPED;
      } else {
        ds = NWB_ADVANCED;
#pragma unroll
        for (int i = 0; i < 5; ++i) q[16 + i] = ll[i];
      }
    }
  }
  ds = __shfl_sync(FULL, ds, 0);
  __syncwarp();
  if (ds == NWB_TRAPPED) return NWB_TRAPPED;
  return ds == NWB_ADVANCED ? NWB_ADVANCED : st;
}

// ---- Philox4x32-10 counter-based stream (DESIGN.md "random streams") ---------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
constexpr uint32_t kStreamSampler = 1, kStreamPlanner = 2;
__device__ __forceinline__ double u01(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }

}  // namespace nwb
