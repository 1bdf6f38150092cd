// nwb_tables.h - host-side construction of the constant tables the validity kernels consume
// (pair runs from the SRDF-derived pair list, scaled foot polygons, scene boxes).  Header-only so the
// flop-counting tool can build the same tables without the CUDA runtime.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "nwb_internal.h"

namespace nwb {

inline void fill_pairs(const nwb_params& params, ValidityTables& t, bool group_only) {
  const int n = params.srdf_trim ? nwbm::kNumPairsTrim : nwbm::kNumPairsNoTrim;
  const unsigned char(*P)[3] = params.srdf_trim ? nwbm::kPairsTrim : nwbm::kPairsNoTrim;
  auto len2 = [](int g) {
    const nwbm::ProxyDef& p = nwbm::kProxies[g];
    double dx = p.b[0] - p.a[0], dy = p.b[1] - p.a[1], dz = p.b[2] - p.a[2];
    return dx * dx + dy * dy + dz * dz;
  };
  int n_entries = 0, n_runs = 0;
  int counts[3] = {0, 0, 0};
  for (int type = 0; type < 3; ++type) {   // 0 = capsule-capsule, 1 = capsule-sphere, 2 = sphere-sphere
    std::map<int, std::vector<int>> by_a;  // first proxy -> partners
    std::vector<int> order;
    for (int k = 0; k < n; ++k) {
      if (group_only && !P[k][2]) continue;
      int a = P[k][0], b = P[k][1];
      int sa = nwbm::kProxies[a].is_sphere, sb = nwbm::kProxies[b].is_sphere;
      int ty = sa + sb;
      if (ty != type) continue;
      if (ty == 1 && sa) std::swap(a, b);   // capsule first
      if (!by_a.count(a)) order.push_back(a);
      by_a[a].push_back(b);
    }
    for (int a : order) {
      PairRun& run = t.runs[n_runs++];
      run.off_a = nwbm::kGeomSlot[a];
      double la = len2(a);
      run.a = (float)la;
      run.inv_a = la > 0 ? (float)(1.0 / la) : 0.f;
      run.first = n_entries;
      run.count = (int)by_a[a].size();
      for (int b : by_a[a]) {
        PairEntry& e = t.entries[n_entries++];
        e.off_b = nwbm::kGeomSlot[b];
        double lb = len2(b);
        e.e = (float)lb;
        e.inv_e = lb > 0 ? (float)(1.0 / lb) : 0.f;
        e.rsum = (float)(nwbm::kProxies[a].r + nwbm::kProxies[b].r);
      }
      counts[type]++;
    }
  }
  t.n_runs_cc = counts[0];
  t.n_runs_cs = counts[1];
  t.n_runs_ss = counts[2];
  // robot proxies tested against the world: the group's links (everything but torso / head) for
  // isFeasible (CollisionRequest.group_name, RP:1250-1251), the whole robot for isPathValid
  t.n_env_links = 0;
  for (int g = 0; g < nwbm::kNumGeom; ++g) {
    if (group_only && g < 3) continue;
    EnvLink& el = t.env_links[t.n_env_links++];
    el.off = nwbm::kGeomSlot[g];
    double l = len2(g);
    el.inv_a = l > 0 ? (float)(1.0 / l) : 0.f;
    el.radius = (float)nwbm::kProxies[g].r;
    el.is_sphere = nwbm::kProxies[g].is_sphere;
    el.half_len = (float)(0.5 * std::sqrt(l));
  }
}

// which = 0: isFeasible pair set (group filter); 1: isPathValid pair set (whole robot)
inline void build_validity_tables(const nwb_params& params, const std::vector<SceneBox>& boxes, bool group_only,
                                  ValidityTables& t) {
  std::memset(&t, 0, sizeof(t));
  // TestStability::scaleConvexHull: scale about the sole origin (0), the foot centre of the model (1: mean of the
  // foot-mesh vertices, calibrated - see the model file) or the mean of the polygon vertices (2)
  double cx = 0, cy = 0;
  if (params.scale_origin == 1) {
    cx = nwbm::kFootCenter[0];
    cy = nwbm::kFootCenter[1];
  } else if (params.scale_origin == 2) {
    for (int i = 0; i < nwbm::kFootN; ++i) {
      cx += nwbm::kFootRight[i][0];
      cy += nwbm::kFootRight[i][1];
    }
    cx /= nwbm::kFootN;
    cy /= nwbm::kFootN;
  }
  const double s = params.scale_sp;
  for (int i = 0; i < nwbm::kFootN; ++i) {
    double x = cx + s * (nwbm::kFootRight[i][0] - cx), y = cy + s * (nwbm::kFootRight[i][1] - cy);
    t.foot_right[i][0] = (float)x;
    t.foot_right[i][1] = (float)y;
    t.foot_left[i][0] = (float)x;
    t.foot_left[i][1] = (float)(-y);
  }
  fill_pairs(params, t, group_only);
  // up to kMaxBoxes boxes travel inline in the kernel parameters; larger scenes go through t.scene (set by the caller)
  t.n_boxes = boxes.size() <= (size_t)kMaxBoxes ? (int)boxes.size() : 0;
  for (int b = 0; b < t.n_boxes; ++b) t.boxes[b] = boxes[b];
  t.env_cull = 1;
}

// Axis-aligned bounds of an oriented box.
inline void box_aabb(const SceneBox& b, float* lo, float* hi) {
  for (int k = 0; k < 3; ++k) {
    const float e = std::fabs(b.r[3 * k + 0]) * b.h[0] + std::fabs(b.r[3 * k + 1]) * b.h[1] + std::fabs(b.r[3 * k + 2]) * b.h[2];
    lo[k] = b.c[k] - e;
    hi[k] = b.c[k] + e;
  }
}

// Flattened bounding-volume hierarchy over the scene boxes: median split of the box centres along the widest axis,
// one box per leaf, nodes in depth-first order with skip links.
inline void build_scene_bvh(const std::vector<SceneBox>& boxes, std::vector<SceneNode>& nodes) {
  nodes.clear();
  const int n = (int)boxes.size();
  if (n == 0) return;
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  nodes.reserve(2 * n);
  struct Rec {
    static int build(const std::vector<SceneBox>& boxes, std::vector<int>& order, int first, int last, std::vector<SceneNode>& nodes) {
      const int me = (int)nodes.size();
      nodes.push_back(SceneNode{});
      float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f}, clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
      for (int i = first; i < last; ++i) {
        float l[3], h[3];
        box_aabb(boxes[order[i]], l, h);
        for (int k = 0; k < 3; ++k) {
          lo[k] = std::min(lo[k], l[k]);
          hi[k] = std::max(hi[k], h[k]);
          clo[k] = std::min(clo[k], boxes[order[i]].c[k]);
          chi[k] = std::max(chi[k], boxes[order[i]].c[k]);
        }
      }
      if (last - first == 1) {
        nodes[me].box = order[first];
      } else {
        nodes[me].box = -1;
        int ax = 0;
        for (int k = 1; k < 3; ++k)
          if (chi[k] - clo[k] > chi[ax] - clo[ax]) ax = k;
        const int mid = (first + last) / 2;
        std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + last,
                         [&](int a, int b) { return boxes[a].c[ax] < boxes[b].c[ax] || (boxes[a].c[ax] == boxes[b].c[ax] && a < b); });
        build(boxes, order, first, mid, nodes);
        build(boxes, order, mid, last, nodes);
      }
      for (int k = 0; k < 3; ++k) {
        nodes[me].lo[k] = lo[k];
        nodes[me].hi[k] = hi[k];
      }
      nodes[me].skip = (int)nodes.size();
      return me;
    }
  };
  Rec::build(boxes, order, 0, n, nodes);
}

}  // namespace nwb
