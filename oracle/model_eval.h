// TEST INFRASTRUCTURE - CPU oracle, never shipped, never on the product path.
//
// fp64 restatement of what RRT_Planner::isFeasible evaluates
// (rrt_connect_planner/src/rrt_planner.cpp:1219-1287; identical copy
// rrt_connect_planner/src/stable_config_generator.cpp:77-159):
//   * MoveIt RobotState forward kinematics over the robot tree  (called rrt_planner.cpp:715-720)
//   * PlanningScene::checkCollision with the SRDF allowed-collision matrix (rrt_planner.cpp:1250-1257)
//     - the FCL mesh narrowphase is replaced by the capsule proxies of the authored model;
//       bit-exactness of the product is defined against THESE proxies
//   * hrl_kinematics::Kinematics::computeCOM + TestStability::isPoseStable(SUPPORT_DOUBLE)
//     (called rrt_planner.cpp:1275) restated per SURVEY.md Appendix B.4.
// MoveIt, FCL and hrl_kinematics are third-party, not vendored and not installed: parity for the
// collision geometry and mass distribution is pinned one-sidedly only (every configuration the
// reference shipped as valid must stay valid - tests/test_oracle_golden.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "../nao_wholebody_planning_b200/csrc/nao_model_gen.h"
#include "kdl_restated.h"

namespace oracle {
using kdlr::Frame;
using kdlr::Rot;
using kdlr::Vec;

struct Box {  // oriented box of the planning scene, expressed in the r_sole frame
  Vec c;      // centre
  Rot R;      // orientation (columns = box axes)
  Vec h;      // half extents
};

struct EvalParams {
  double scale_sp = nwbm::kSupportScaleDefault;
  int scale_origin = nwbm::kScaleOriginDefault;
  int srdf_trim = 1;
  std::vector<Box> boxes;
};

// Generic table-driven FK: frame = parent * Trans(pre) * Rot(axis, sign*q[joint]) * Trans(post).
inline void fk_all(const double* q, Frame* out /*[kNumFrames]*/) {
  for (int i = 0; i < nwbm::kNumFrames; ++i) {
    const nwbm::FrameDef& d = nwbm::kFrames[i];
    if (d.parent < 0) {
      out[i] = Frame{};
      continue;
    }
    Frame f = out[d.parent];
    f.p = f.p + f.M * Vec{d.pre[0], d.pre[1], d.pre[2]};
    if (d.axis_kind != nwbm::AX_NONE && d.joint >= 0) {
      double a = d.sign * q[d.joint];
      Rot r;
      switch (d.axis_kind) {
        case nwbm::AX_X: r = kdlr::rot_x(a); break;
        case nwbm::AX_Y: r = kdlr::rot_y(a); break;
        case nwbm::AX_Z: r = kdlr::rot_z(a); break;
        default: r = kdlr::rot_axis(Vec{d.axis[0], d.axis[1], d.axis[2]}, a); break;
      }
      f.M = f.M * r;
    }
    f.p = f.p + f.M * Vec{d.post[0], d.post[1], d.post[2]};
    out[i] = f;
  }
}

// hrl_kinematics computeCOM: com = sum m_i (T_i * cog_i) / sum m_i.
inline Vec center_of_mass(const Frame* fr) {
  Vec acc;
  double tot = 0;
  for (int i = 0; i < nwbm::kNumMass; ++i) {
    const nwbm::MassDef& m = nwbm::kMasses[i];
    const Frame& f = fr[m.frame];
    acc = acc + (f.p + f.M * Vec{m.com[0], m.com[1], m.com[2]}) * m.mass;
    tot += m.mass;
  }
  return acc * (1.0 / tot);
}

struct P2 {
  double x, y;
};

// TestStability::scaleConvexHull: both foot polygons scaled by s about the foot centre (scale_origin 1: the
// model's calibrated mean of the foot-mesh vertices, models/nao_h25_rfoot.json foot_center_doc), about the sole
// origin (0) or about the mean of the polygon vertices (2) - SURVEY.md Appendix B.4.  Left = right mirrored in y.
inline void foot_polygons(const EvalParams& ep, std::vector<P2>& right, std::vector<P2>& left) {
  const int n = nwbm::kFootN;
  right.resize(n);
  left.resize(n);
  double cx = 0, cy = 0;
  if (ep.scale_origin == 1) {
    cx = nwbm::kFootCenter[0];
    cy = nwbm::kFootCenter[1];
  } else if (ep.scale_origin == 2) {
    for (int i = 0; i < n; ++i) {
      cx += nwbm::kFootRight[i][0] / n;
      cy += nwbm::kFootRight[i][1] / n;
    }
  }
  for (int i = 0; i < n; ++i) {
    double x = nwbm::kFootRight[i][0], y = nwbm::kFootRight[i][1];
    right[i] = {cx + ep.scale_sp * (x - cx), cy + ep.scale_sp * (y - cy)};
    left[i] = {cx + ep.scale_sp * (x - cx), -(cy + ep.scale_sp * (y - cy))};
  }
}

// Andrew monotone chain; returns the hull CLOCKWISE (as the logged hull of
// rrt_connect_planner_example/src/polygon_log.txt is), collinear points dropped.
inline std::vector<P2> convex_hull(std::vector<P2> pts) {
  std::sort(pts.begin(), pts.end(), [](const P2& a, const P2& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
  pts.erase(std::unique(pts.begin(), pts.end(), [](const P2& a, const P2& b) { return a.x == b.x && a.y == b.y; }),
            pts.end());
  const int n = (int)pts.size();
  if (n < 3) return pts;
  auto cr = [](const P2& o, const P2& a, const P2& b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); };
  std::vector<P2> h(2 * n);
  int k = 0;
  for (int i = 0; i < n; ++i) {  // lower hull
    while (k >= 2 && cr(h[k - 2], h[k - 1], pts[i]) <= 0) --k;
    h[k++] = pts[i];
  }
  for (int i = n - 2, t = k + 1; i >= 0; --i) {  // upper hull
    while (k >= t && cr(h[k - 2], h[k - 1], pts[i]) <= 0) --k;
    h[k++] = pts[i];
  }
  h.resize(k - 1);                  // counter-clockwise, starts at the min-x point
  std::reverse(h.begin() + 1, h.end());  // -> clockwise, still starting at the min-x point
  return h;
}

// TestStability::pointInConvexHull: the first edge fixes the expected
// sign; any edge whose strict sign differs -> outside.
inline bool point_in_hull(const P2& p, const std::vector<P2>& poly) {
  bool positive = false;
  for (size_t i = 0; i < poly.size(); ++i) {
    size_t i2 = (i + 1) % poly.size();
    double dx = poly[i2].x - poly[i].x, dy = poly[i2].y - poly[i].y;
    if (dx == 0.0 && dy == 0.0) continue;
    double t = (p.y - poly[i].y) * dx - (p.x - poly[i].x) * dy;
    if (i == 0) positive = (t > 0.0);
    if ((t > 0.0) != positive) return false;
  }
  return true;
}

// Signed distance from p to the hull boundary: > 0 inside, < 0 outside (exact for a convex polygon).
inline double hull_margin(const P2& p, const std::vector<P2>& poly) {
  // orientation-agnostic: use the polygon's signed area to orient the edge normals
  double area = 0;
  for (size_t i = 0; i < poly.size(); ++i) {
    size_t j = (i + 1) % poly.size();
    area += poly[i].x * poly[j].y - poly[j].x * poly[i].y;
  }
  double sgn = area >= 0 ? 1.0 : -1.0;
  double inside = 1e300;
  bool in = true;
  double dmin2 = 1e300;
  for (size_t i = 0; i < poly.size(); ++i) {
    size_t j = (i + 1) % poly.size();
    double dx = poly[j].x - poly[i].x, dy = poly[j].y - poly[i].y;
    double len2 = dx * dx + dy * dy;
    if (len2 == 0) continue;
    double t = sgn * (dx * (p.y - poly[i].y) - dy * (p.x - poly[i].x)) / std::sqrt(len2);
    if (t <= 0) in = false;
    inside = std::min(inside, t);
    double u = ((p.x - poly[i].x) * dx + (p.y - poly[i].y) * dy) / len2;
    u = std::min(1.0, std::max(0.0, u));
    double ex = poly[i].x + u * dx - p.x, ey = poly[i].y + u * dy - p.y;
    dmin2 = std::min(dmin2, ex * ex + ey * ey);
  }
  return in ? inside : -std::sqrt(dmin2);
}

// Support polygon of TestStability::isPoseStable(SUPPORT_DOUBLE): right-foot polygon united with
// the left-foot polygon carried into the right-sole frame, plane normal (0,0,1) -> drop z.
inline std::vector<P2> support_hull(const EvalParams& ep, const Frame* fr) {
  std::vector<P2> right, left;
  foot_polygons(ep, right, left);
  const Frame& rs = fr[nwbm::kFrameRSole];
  const Frame& ls = fr[nwbm::kFrameLSole];
  Rot Rt = kdlr::transpose(rs.M);
  std::vector<P2> pts = right;
  for (auto& v : left) {
    Vec w = ls.p + ls.M * Vec{v.x, v.y, 0};
    Vec l = Rt * (w - rs.p);
    pts.push_back({l.x, l.y});
  }
  return convex_hull(pts);
}

inline bool is_pose_stable(const EvalParams& ep, const Frame* fr, double* margin) {
  std::vector<P2> hull = support_hull(ep, fr);
  if (hull.size() < 3) {
    if (margin) *margin = -1e30;
    return false;
  }
  const Frame& rs = fr[nwbm::kFrameRSole];
  Vec c = kdlr::transpose(rs.M) * (center_of_mass(fr) - rs.p);
  P2 p{c.x, c.y};
  if (margin) *margin = hull_margin(p, hull);
  return point_in_hull(p, hull);
}

// Closest distance between two segments (Ericson, "Real-Time Collision Detection" 5.1.9).
inline double seg_seg_distance(Vec p1, Vec q1, Vec p2, Vec q2) {
  Vec d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  double a = kdlr::dot(d1, d1), e = kdlr::dot(d2, d2), f = kdlr::dot(d2, r);
  double s, t;
  const double EPS = 1e-18;
  if (a <= EPS && e <= EPS) {
    s = t = 0;
  } else if (a <= EPS) {
    s = 0;
    t = std::min(1.0, std::max(0.0, f / e));
  } else {
    double c = kdlr::dot(d1, r);
    if (e <= EPS) {
      t = 0;
      s = std::min(1.0, std::max(0.0, -c / a));
    } else {
      double b = kdlr::dot(d1, d2);
      double denom = a * e - b * b;
      s = denom > EPS ? std::min(1.0, std::max(0.0, (b * f - c * e) / denom)) : 0.0;
      t = (b * s + f) / e;
      if (t < 0) {
        t = 0;
        s = std::min(1.0, std::max(0.0, -c / a));
      } else if (t > 1) {
        t = 1;
        s = std::min(1.0, std::max(0.0, (b - c) / a));
      }
    }
  }
  Vec d = r + d1 * s - d2 * t;
  return std::sqrt(kdlr::dot(d, d));
}

// Squared distance from point p to an oriented box.
inline double point_box_dist2(Vec p, const Box& b) {
  Vec l = kdlr::transpose(b.R) * (p - b.c);
  double dx = std::max(std::fabs(l.x) - b.h.x, 0.0), dy = std::max(std::fabs(l.y) - b.h.y, 0.0),
         dz = std::max(std::fabs(l.z) - b.h.z, 0.0);
  return dx * dx + dy * dy + dz * dz;
}

// Exact distance from segment a-b to an oriented box.  In the box frame
// f(t) = sum_k max(|x_k(t)| - h_k, 0)^2 with x(t) = a + t d is convex and piecewise quadratic in t;
// its break points are the (at most six) face crossings.  Between consecutive break points f is a
// single parabola, so the minimum is the best clamped vertex over the pieces.
inline double seg_box_distance(Vec a, Vec b, const Box& box) {
  Rot Rt = kdlr::transpose(box.R);
  Vec la = Rt * (a - box.c), ld = Rt * (b - a);
  const double A[3] = {la.x, la.y, la.z}, D[3] = {ld.x, ld.y, ld.z}, H[3] = {box.h.x, box.h.y, box.h.z};
  std::vector<double> ts = {0.0, 1.0};
  for (int k = 0; k < 3; ++k)
    if (D[k] != 0.0)
      for (double s : {-1.0, 1.0}) {
        double t = (s * H[k] - A[k]) / D[k];
        if (t > 0.0 && t < 1.0) ts.push_back(t);
      }
  std::sort(ts.begin(), ts.end());
  auto f_at = [&](double t) {
    double acc = 0;
    for (int k = 0; k < 3; ++k) {
      double g = std::max(std::fabs(A[k] + t * D[k]) - H[k], 0.0);
      acc += g * g;
    }
    return acc;
  };
  double best = std::min(f_at(0.0), f_at(1.0));
  for (size_t i = 0; i + 1 < ts.size(); ++i) {
    double t0 = ts[i], t1 = ts[i + 1];
    if (t1 <= t0) continue;
    double mid = 0.5 * (t0 + t1), qa = 0, qb = 0;  // f = qa t^2 + qb t + qc on this piece
    for (int k = 0; k < 3; ++k) {
      double x = A[k] + mid * D[k];
      if (std::fabs(x) <= H[k]) continue;
      double sg = x > 0 ? 1.0 : -1.0;
      double c0 = sg * A[k] - H[k], c1 = sg * D[k];
      qa += c1 * c1;
      qb += 2 * c0 * c1;
    }
    double tv = qa > 0 ? -qb / (2 * qa) : t0;
    tv = std::min(t1, std::max(t0, tv));
    best = std::min(best, std::min(f_at(tv), f_at(t1)));
  }
  return std::sqrt(best);
}

struct PairTable {
  int n;
  const unsigned char (*p)[3];
};
inline PairTable pair_table(int srdf_trim) {
  if (srdf_trim) return {nwbm::kNumPairsTrim, nwbm::kPairsTrim};
  return {nwbm::kNumPairsNoTrim, nwbm::kPairsNoTrim};
}

// PlanningScene::checkCollision restated on the capsule proxies.  group_only applies the
// CollisionRequest.group_name filter (rrt_planner.cpp:1250-1251); isPathValid uses the whole robot.
// Returns true when in collision; *min_sep = min over tested pairs of (distance - r_a - r_b).
inline bool in_collision(const EvalParams& ep, const Frame* fr, bool group_only, double* min_sep) {
  Vec A[nwbm::kNumGeom], B[nwbm::kNumGeom];
  for (int i = 0; i < nwbm::kNumGeom; ++i) {
    const nwbm::ProxyDef& pr = nwbm::kProxies[i];
    const Frame& f = fr[pr.frame];
    A[i] = f.p + f.M * Vec{pr.a[0], pr.a[1], pr.a[2]};
    B[i] = f.p + f.M * Vec{pr.b[0], pr.b[1], pr.b[2]};
  }
  PairTable pt = pair_table(ep.srdf_trim);
  double best = 1e300;
  for (int k = 0; k < pt.n; ++k) {
    if (group_only && !pt.p[k][2]) continue;
    int a = pt.p[k][0], b = pt.p[k][1];
    double d = seg_seg_distance(A[a], B[a], A[b], B[b]) - nwbm::kProxies[a].r - nwbm::kProxies[b].r;
    best = std::min(best, d);
  }
  // robot vs world: every geometry link that belongs to the group (all but torso/head; the
  // whole robot for isPathValid) against every scene box.
  for (size_t bi = 0; bi < ep.boxes.size(); ++bi)
    for (int i = 0; i < nwbm::kNumGeom; ++i) {
      if (group_only && i < 3) continue;  // torso, HeadYaw_link, HeadPitch_link are not in the group
      double d = seg_box_distance(A[i], B[i], ep.boxes[bi]) - nwbm::kProxies[i].r;
      best = std::min(best, d);
    }
  if (min_sep) *min_sep = best;
  return best < 0.0;
}

// RRT_Planner::isFeasible: both tests always run (no early exit, rrt_planner.cpp:1257-1286).
// reason bit0 = collision, bit1 = unstable.  margins = {min separation, hull margin}.
inline bool is_feasible(const EvalParams& ep, const double* q, unsigned* reason, double* margins) {
  Frame fr[nwbm::kNumFrames];
  fk_all(q, fr);
  double sep, hm;
  bool coll = in_collision(ep, fr, /*group_only=*/true, &sep);
  bool stable = is_pose_stable(ep, fr, &hm);
  if (reason) *reason = (coll ? 1u : 0u) | (stable ? 0u : 2u);
  if (margins) {
    margins[0] = sep;
    margins[1] = hm;
  }
  return !coll && stable;
}

}  // namespace oracle
