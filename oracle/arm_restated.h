// TEST INFRASTRUCTURE - CPU oracle, never shipped, never on the product path.
//
// fp64 restatement of the right-arm pieces of the reference planner:
//   NCK = rrt_connect_planner/src/nao_constraint_kinematics.cpp
//   AOC = rrt_connect_planner/src/articulated_object_constraints.cpp
//   IKF = rrt_connect_planner/src/ikfast_TranslationDirection5D.cpp   (generated, ikfast v55)
//
// The reference solves "hand position + hand z-axis direction" for the 5 right-arm joints with the
// generated solver IKF:341-1517 (a degree-8 polynomial root search).  That file is NOT restated
// line by line: the arm it describes (IKF:301-326, its fk) decouples, so the same solution set is
// obtained in closed form -
//     the forearm axis x_f is perpendicular to the requested direction d      -> one angle phi
//     |elbow - shoulder| = |upper arm|                                         -> A cos phi + B sin phi = C
//     shoulder pitch/roll from the elbow position, elbow yaw/roll from x_f, wrist yaw from d
// giving 2 x 2 x 2 = 8 solutions.  PINNED by tests/test_arm_ik_cpu.py against (i) the reference
// solver itself, compiled where it lies into oracle/_ref/libikfast_ref.so (same solution sets to
// 1e-9 on seeded targets), (ii) tests/golden/ikfast_vectors.npz generated from it, and (iii) the
// shipped goal-configuration files: their right-arm columns are one of
// our solutions for the hand pose the file was generated for.
#pragma once
#include <cmath>

#include "kdl_restated.h"

namespace oracle {

// The solver's arm model, read off its forward kinematics (IKF:301-326); torso frame.
constexpr double kArmShoulder[3] = {0.0, -0.098, 0.100};
constexpr double kArmUpperX = 0.105, kArmUpperY = -0.015;   // elbow in the shoulder-roll frame
constexpr double kArmForearm = 0.12395;                     // elbow -> wrist along the forearm axis
constexpr double kArmHandOffset = 0.0159;                   // wrist -> grasp point, perpendicular to x_f and d

// position and direction of the grasp frame: S + Ry(q0) Rz(q1) [a + Rx(q2) Rz(q3) (b + Rx(q4) c0)],
// d = Ry Rz Rx Rz Rx (0,-1,0), c0 = (0,0,-0.0159).
inline void arm5d_fk(const double* q, double* pos, double* dir) {
  using namespace kdlr;
  Rot R01 = rot_y(q[0]) * rot_z(q[1]);
  Rot R03 = R01 * rot_x(q[2]) * rot_z(q[3]);
  Rot R04 = R03 * rot_x(q[4]);
  Vec p = Vec{kArmShoulder[0], kArmShoulder[1], kArmShoulder[2]} + R01 * Vec{kArmUpperX, kArmUpperY, 0} +
          R03 * Vec{kArmForearm, 0, 0} + R04 * Vec{0, 0, -kArmHandOffset};
  Vec d = R04 * Vec{0, -1, 0};
  pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
  dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
}

// All solutions of arm5d_fk(q) = (pos, dir), each joint in (-pi, pi].  `dir` must be a unit vector.
// Returns the number of solutions written to out[8][5] (0, 4 or 8; fewer at singular poses).
inline int arm5d_ik(const double* pos, const double* dir, double* out /*[8][5]*/) {
  using namespace kdlr;
  const double L = kArmForearm, h = kArmHandOffset;
  const double rho2 = kArmUpperX * kArmUpperX + kArmUpperY * kArmUpperY, rho = std::sqrt(rho2);
  const double delta = std::atan2(-kArmUpperY, kArmUpperX);  // a = rho (cos delta, -sin delta, 0)
  Vec d{dir[0], dir[1], dir[2]};
  Vec r{pos[0] - kArmShoulder[0], pos[1] - kArmShoulder[1], pos[2] - kArmShoulder[2]};
  // orthonormal (u, v, d), u x v = d
  Vec t = std::fabs(d.x) < 0.9 ? Vec{1, 0, 0} : Vec{0, 1, 0};
  Vec u = cross(t, d);
  u = u * (1.0 / std::sqrt(dot(u, u)));
  Vec v = cross(d, u);
  const double ru = dot(r, u), rv = dot(r, v);
  const double A = -L * ru + h * rv, B = -L * rv - h * ru;
  const double C = 0.5 * (rho2 - dot(r, r) - L * L - h * h);
  const double amp = std::sqrt(A * A + B * B);
  if (!(amp > 0) || std::fabs(C) > amp) return 0;
  const double phi0 = std::atan2(B, A), dphi = std::acos(C / amp);
  int n = 0;
  for (int s_phi = 0; s_phi < 2; ++s_phi) {
    const double phi = phi0 + (s_phi ? -dphi : dphi);
    const double cp = std::cos(phi), sp = std::sin(phi);
    Vec xf = u * cp + v * sp;                                   // forearm axis
    Vec e = r - xf * L + cross(d, xf) * h;                      // elbow - shoulder
    double sy = e.y / rho;
    if (sy > 1) sy = 1;
    if (sy < -1) sy = -1;
    for (int s_sh = 0; s_sh < 2; ++s_sh) {
      const double a1 = std::asin(sy);
      double q1 = delta + (s_sh ? M_PI - a1 : a1);
      const double radial = rho * std::cos(q1 - delta);         // 0.105 c1 + 0.015 s1
      if (std::fabs(radial) < 1e-12) continue;                  // elbow on the shoulder pitch axis
      const double sg = radial > 0 ? 1.0 : -1.0;
      const double q0 = std::atan2(-e.z * sg, e.x * sg);
      Rot R01 = rot_y(q0) * rot_z(q1);
      Vec f = transpose(R01) * xf;                              // (c3, c2 s3, s2 s3)
      double c3 = f.x;
      if (c3 > 1) c3 = 1;
      if (c3 < -1) c3 = -1;
      for (int s_el = 0; s_el < 2; ++s_el) {
        const double q3 = s_el ? -std::acos(c3) : std::acos(c3);
        const double s3 = std::sin(q3);
        if (std::fabs(s3) < 1e-12) continue;                    // straight elbow: yaw indeterminate
        const double q2 = std::atan2(f.z / s3, f.y / s3);
        Rot R03 = R01 * rot_x(q2) * rot_z(q3);
        Vec dl = transpose(R03) * d;                            // Rx(q4) (0,-1,0) = (0, -c4, -s4)
        const double q4 = std::atan2(-dl.z, -dl.y);
        double* o = out + n * 5;
        o[0] = q0; o[1] = q1; o[2] = q2; o[3] = q3; o[4] = q4;
        for (int j = 0; j < 5; ++j) {
          if (o[j] > M_PI) o[j] -= 2 * M_PI;
          if (o[j] <= -M_PI) o[j] += 2 * M_PI;
        }
        ++n;
      }
    }
  }
  return n;
}

}  // namespace oracle
