// TEST INFRASTRUCTURE - CPU oracle, never shipped, never on the product path.
//
// Restatement (fp64) of the orocos-KDL pieces the reference calls on its hot path.  KDL itself is
// a third-party dependency that is NOT vendored in the reference (package.xml lists `orocos_kdl`
// with no version; API vintage 1.1-1.3) and not installed here, so its published algorithms are
// restated from their documented behaviour and anchored on the
// reference's own call sites and data files:
//   chains          rrt_connect_planner/src/local_planner.cpp:18-47,
//                   rrt_connect_planner/src/nao_constraint_kinematics.cpp:23-47
//   FK solver       ChainFkSolverPos_recursive  - constructed at local_planner.cpp:120,180,251
//   Jacobian        ChainJntToJacSolver         - used inside ChainIkSolverVel_pinv
//   velocity IK     ChainIkSolverVel_pinv(eps=1e-5, maxiter=150) - local_planner.cpp:181
//   position IK     ChainIkSolverPos_NR(maxiter=100, eps=1e-3)   - local_planner.cpp:182
// Pinned by tests/test_oracle_golden.py: the 463 shipped configurations reproduce through this
// code to file precision.
#pragma once
#include <cmath>
#include <utility>
#include <vector>

namespace kdlr {

struct Vec {
  double x = 0, y = 0, z = 0;
};
inline Vec operator+(Vec a, Vec b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec operator-(Vec a, Vec b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec operator*(Vec a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(Vec a, Vec b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec cross(Vec a, Vec b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

struct Rot {  // row-major 3x3, m[3*r+c]
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double operator()(int r, int c) const { return m[3 * r + c]; }
};
inline Rot operator*(const Rot& a, const Rot& b) {
  Rot o;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) o.m[3 * r + c] = a(r, 0) * b(0, c) + a(r, 1) * b(1, c) + a(r, 2) * b(2, c);
  return o;
}
inline Vec operator*(const Rot& a, Vec v) {
  return {a(0, 0) * v.x + a(0, 1) * v.y + a(0, 2) * v.z, a(1, 0) * v.x + a(1, 1) * v.y + a(1, 2) * v.z,
          a(2, 0) * v.x + a(2, 1) * v.y + a(2, 2) * v.z};
}
inline Rot transpose(const Rot& a) {
  Rot o;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) o.m[3 * r + c] = a(c, r);
  return o;
}
inline Rot rot_x(double a) {
  Rot o;
  double c = std::cos(a), s = std::sin(a);
  o.m[4] = c; o.m[5] = -s; o.m[7] = s; o.m[8] = c;
  return o;
}
inline Rot rot_y(double a) {
  Rot o;
  double c = std::cos(a), s = std::sin(a);
  o.m[0] = c; o.m[2] = s; o.m[6] = -s; o.m[8] = c;
  return o;
}
inline Rot rot_z(double a) {
  Rot o;
  double c = std::cos(a), s = std::sin(a);
  o.m[0] = c; o.m[1] = -s; o.m[3] = s; o.m[4] = c;
  return o;
}
// Rotation about a unit axis (Rodrigues), KDL::Rotation::Rot2.
inline Rot rot_axis(Vec k, double a) {
  double c = std::cos(a), s = std::sin(a), v = 1 - c;
  Rot o;
  o.m[0] = k.x * k.x * v + c;       o.m[1] = k.x * k.y * v - k.z * s; o.m[2] = k.x * k.z * v + k.y * s;
  o.m[3] = k.y * k.x * v + k.z * s; o.m[4] = k.y * k.y * v + c;       o.m[5] = k.y * k.z * v - k.x * s;
  o.m[6] = k.z * k.x * v - k.y * s; o.m[7] = k.z * k.y * v + k.x * s; o.m[8] = k.z * k.z * v + c;
  return o;
}
// KDL::Rotation::RPY(roll, pitch, yaw) = RotZ(yaw) * RotY(pitch) * RotX(roll)
inline Rot rot_rpy(double r, double p, double y) { return rot_z(y) * rot_y(p) * rot_x(r); }

struct Frame {
  Rot M;
  Vec p;
};
inline Frame operator*(const Frame& a, const Frame& b) { return {a.M * b.M, a.M * b.p + a.p}; }

enum JointType { None = 0, RotX = 1, RotY = 2, RotZ = 3 };

// KDL::Segment(Joint(type), Frame f_tip): pose(q) = joint.pose(q) * f_tip  (joint first, then tip).
struct Segment {
  JointType joint;
  Frame tip;
  Frame pose(double q) const {
    Frame j;
    if (joint == RotX) j.M = rot_x(q);
    else if (joint == RotY) j.M = rot_y(q);
    else if (joint == RotZ) j.M = rot_z(q);
    return j * tip;
  }
  Vec axis() const {
    if (joint == RotX) return {1, 0, 0};
    if (joint == RotY) return {0, 1, 0};
    return {0, 0, 1};
  }
};

struct Chain {
  std::vector<Segment> segs;
  void add(JointType j, Vec tip) { segs.push_back({j, Frame{Rot{}, tip}}); }
  void add(JointType j, const Frame& tip) { segs.push_back({j, tip}); }
  int nj() const {
    int n = 0;
    for (auto& s : segs) n += (s.joint != None);
    return n;
  }
};

// ChainFkSolverPos_recursive::JntToCart: out = prod seg_i.pose(q_i); None joints consume no q.
inline Frame fk(const Chain& c, const double* q) {
  Frame f;
  int j = 0;
  for (auto& s : c.segs) {
    double qi = 0;
    if (s.joint != None) qi = q[j++];
    f = f * s.pose(qi);
  }
  return f;
}

// ChainJntToJacSolver::JntToJac: 6 x nj geometric Jacobian in the base frame, reference point at
// the chain tip: column i = (a_i x (p_tip - o_i), a_i).  J is row-major [6][nj].
inline void jacobian(const Chain& c, const double* q, double* J) {
  const int n = c.nj();
  Frame f;
  std::vector<Vec> origin(n), axis(n);
  int j = 0;
  for (auto& s : c.segs) {
    double qi = 0;
    if (s.joint != None) {
      qi = q[j];
      origin[j] = f.p;            // joint rotates about the start of its segment
      axis[j] = f.M * s.axis();
      ++j;
    }
    f = f * s.pose(qi);
  }
  for (int i = 0; i < n; ++i) {
    Vec v = cross(axis[i], f.p - origin[i]);
    J[0 * n + i] = v.x; J[1 * n + i] = v.y; J[2 * n + i] = v.z;
    J[3 * n + i] = axis[i].x; J[4 * n + i] = axis[i].y; J[5 * n + i] = axis[i].z;
  }
}

// KDL::Rotation::GetRot(): rotation vector axis*angle (GetRotAngle with epsilon = 1e-6).
inline Vec rot_vector(const Rot& R) {
  const double eps = 1e-6;
  double ca = (R.m[0] + R.m[4] + R.m[8] - 1) / 2.0;
  double t = eps * eps / 2.0;
  if (ca > 1 - t) return {0, 0, 0};  // angle 0 (axis arbitrary)
  if (ca < -1 + t) {                 // angle pi
    double x = std::sqrt((R.m[0] + 1.0) / 2), y = std::sqrt((R.m[4] + 1.0) / 2), z = std::sqrt((R.m[8] + 1.0) / 2);
    if (R.m[2] < 0) x = -x;
    if (R.m[7] < 0) y = -y;
    if (x * y * R.m[1] < 0) x = -x;
    return Vec{x, y, z} * M_PI;
  }
  double ax = R.m[7] - R.m[5], ay = R.m[2] - R.m[6], az = R.m[3] - R.m[1];
  double mod = std::sqrt(ax * ax + ay * ay + az * az);
  double angle = std::atan2(mod / 2, ca);
  return Vec{ax / mod, ay / mod, az / mod} * angle;
}

// KDL::diff(F_a, F_b) = Twist(p_b - p_a, R_a * rotvec(R_a^T R_b)); out = {vx,vy,vz,wx,wy,wz}.
inline void diff(const Frame& a, const Frame& b, double* tw) {
  Vec dp = b.p - a.p;
  Vec w = a.M * rot_vector(transpose(a.M) * b.M);
  tw[0] = dp.x; tw[1] = dp.y; tw[2] = dp.z; tw[3] = w.x; tw[4] = w.y; tw[5] = w.z;
}

// Thin SVD of a row-major m x n matrix (m >= n) by one-sided Jacobi (Hestenes): A = U diag(S) V^T.
// KDL uses a Householder SVD (svd_HH, maxiter 150); the decomposition it converges to is the same
// up to column signs / order, which do not affect the pseudo-inverse built from it.
// Column pairs are visited in round-robin order (rounds of disjoint pairs; for n = 5:
// (1,4)(2,3) (0,2)(3,4) (0,4)(1,3) (0,1)(2,4) (0,3)(1,2)) and a pair counts as converged when
// |gamma| <= 1e-15 sqrt(alpha beta) - the specification shared with the product.
inline void svd_jacobi(const double* A, int m, int n, double* U, double* S, double* V) {
  std::vector<double> W(A, A + m * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
  // round-robin tournament over n players (n odd: one sits out per round)
  std::vector<std::pair<int, int>> order;
  if (n == 5) {
    order = {{1, 4}, {2, 3}, {0, 2}, {3, 4}, {0, 4}, {1, 3}, {0, 1}, {2, 4}, {0, 3}, {1, 2}};
  } else {
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) order.push_back({p, q});
  }
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (auto pq : order) {
      const int p = pq.first, q = pq.second;
      double alpha = 0, beta = 0, gamma = 0;
      for (int r = 0; r < m; ++r) {
        alpha += W[r * n + p] * W[r * n + p];
        beta += W[r * n + q] * W[r * n + q];
        gamma += W[r * n + p] * W[r * n + q];
      }
      if (!(gamma * gamma > 1e-30 * (alpha * beta))) continue;
      rotated = true;
      const double d = beta - alpha;
      const double sg = ((d >= 0) == (gamma >= 0) || d == 0) ? 1.0 : -1.0;
      const double t = sg * (2.0 * std::fabs(gamma)) / (std::fabs(d) + std::sqrt(d * d + 4.0 * gamma * gamma));
      const double c = 1 / std::sqrt(1 + t * t), s = c * t;
      for (int r = 0; r < m; ++r) {
        double wp = W[r * n + p], wq = W[r * n + q];
        W[r * n + p] = c * wp - s * wq;
        W[r * n + q] = s * wp + c * wq;
      }
      for (int r = 0; r < n; ++r) {
        double vp = V[r * n + p], vq = V[r * n + q];
        V[r * n + p] = c * vp - s * vq;
        V[r * n + q] = s * vp + c * vq;
      }
    }
    if (!rotated) break;
  }
  for (int j = 0; j < n; ++j) {
    double nrm = 0;
    for (int r = 0; r < m; ++r) nrm += W[r * n + j] * W[r * n + j];
    nrm = std::sqrt(nrm);
    S[j] = nrm;
    for (int r = 0; r < m; ++r) U[r * n + j] = nrm > 0 ? W[r * n + j] / nrm : 0.0;
  }
}

// ChainIkSolverVel_pinv::CartToJnt: qdot = V diag(|s_i| < eps ? 0 : 1/s_i) U^T v.
inline void ik_vel_pinv(const Chain& c, const double* q, const double* twist, double* qdot, double eps = 1e-5) {
  const int n = c.nj();
  std::vector<double> J(6 * n), U(6 * n), S(n), V(n * n), tmp(n);
  jacobian(c, q, J.data());
  svd_jacobi(J.data(), 6, n, U.data(), S.data(), V.data());
  for (int i = 0; i < n; ++i) {
    double sum = 0;
    for (int j = 0; j < 6; ++j) sum += U[j * n + i] * twist[j];
    tmp[i] = sum * (std::fabs(S[i]) < eps ? 0.0 : 1.0 / S[i]);
  }
  for (int i = 0; i < n; ++i) {
    double sum = 0;
    for (int j = 0; j < n; ++j) sum += V[i * n + j] * tmp[j];
    qdot[i] = sum;
  }
}

// ChainIkSolverPos_NR::CartToJnt.  The update is applied BEFORE the convergence test
// (reproduces the shipped database).  Returns 0 on convergence, -3 else;
// *iters_out = number of updates applied.
inline int ik_pos_nr(const Chain& c, const double* q_init, const Frame& target, double* q_out, int maxiter,
                     double eps, double pinv_eps, int* iters_out = nullptr) {
  const int n = c.nj();
  for (int i = 0; i < n; ++i) q_out[i] = q_init[i];
  std::vector<double> dq(n);
  double tw[6];
  int i;
  for (i = 0; i < maxiter; ++i) {
    Frame f = fk(c, q_out);
    diff(f, target, tw);
    ik_vel_pinv(c, q_out, tw, dq.data(), pinv_eps);
    for (int k = 0; k < n; ++k) q_out[k] += dq[k];
    bool conv = true;  // KDL::Equal(delta_twist, Twist::Zero(), eps): every |component| < eps
    for (int k = 0; k < 6; ++k) conv = conv && (eps > tw[k]) && (tw[k] > -eps);
    if (conv) break;
  }
  if (iters_out) *iters_out = (i < maxiter) ? i + 1 : maxiter;
  return (i != maxiter) ? 0 : -3;
}

}  // namespace kdlr
