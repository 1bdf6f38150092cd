// validity_core.cuh - per-configuration validity evaluation, generic over the scalar type T and the
// storage policy of the proxy point table.
//
// Device instantiations: T = float (one configuration per thread, point table in registers or
// shared memory) and T = F2 (two configurations per thread on the packed FP32 instructions, point
// table in shared memory).  Host instantiation: tools/count_flops.cpp runs the very same code with
// a flop-counting scalar to freeze the algorithmic flops per configuration used by the roofline
// (profiles/roofline.json).
#pragma once
#include <type_traits>

#include "nwb_internal.h"
#include "nwb_math.cuh"

namespace nwb {

// Visitor of the generated FK walk: stores proxy world points, accumulates the COM (x,y) and
// remembers the left-sole frame.
template <class T, class Store, bool WANT_COM>
struct ValidityVisitor {
  Store& st;
  T comx, comy;
  XF<T> lsole;
  NWB_HD explicit ValidityVisitor(Store& s) : st(s), comx(T(0.f)), comy(T(0.f)) {}

  template <int ID>
  NWB_HD void frame(const XF<T>& F) {
    using M = nwbm::FrameMass<ID>;
    if constexpr (M::has && WANT_COM) {
      comx = nwb_fma(T(float(M::w)), F.p[0], comx);
      comy = nwb_fma(T(float(M::w)), F.p[1], comy);
      if constexpr (M::wx != 0) { comx = nwb_fma(T(float(M::wx)), F.r[0], comx); comy = nwb_fma(T(float(M::wx)), F.r[3], comy); }
      if constexpr (M::wy != 0) { comx = nwb_fma(T(float(M::wy)), F.r[1], comx); comy = nwb_fma(T(float(M::wy)), F.r[4], comy); }
      if constexpr (M::wz != 0) { comx = nwb_fma(T(float(M::wz)), F.r[2], comx); comy = nwb_fma(T(float(M::wz)), F.r[5], comy); }
    }
    using P = nwbm::FrameProxy<ID>;
    if constexpr (P::has) {
      put<P::slot, (P::ax != 0), (P::ay != 0), (P::az != 0)>(F, T(float(P::ax)), T(float(P::ay)), T(float(P::az)));
      if constexpr (!P::sphere)
        put<P::slot + 3, (P::bx != 0), (P::by != 0), (P::bz != 0)>(F, T(float(P::bx)), T(float(P::by)), T(float(P::bz)));
    }
    if constexpr (ID == nwbm::kFrameLSole) lsole = F;
  }
  template <int SLOT, bool UX, bool UY, bool UZ>
  NWB_HD void put(const XF<T>& F, T x, T y, T z) {
    T wx = F.p[0], wy = F.p[1], wz = F.p[2];
    if constexpr (UX) { wx = nwb_fma(F.r[0], x, wx); wy = nwb_fma(F.r[3], x, wy); wz = nwb_fma(F.r[6], x, wz); }
    if constexpr (UY) { wx = nwb_fma(F.r[1], y, wx); wy = nwb_fma(F.r[4], y, wy); wz = nwb_fma(F.r[7], y, wz); }
    if constexpr (UZ) { wx = nwb_fma(F.r[2], z, wx); wy = nwb_fma(F.r[5], z, wy); wz = nwb_fma(F.r[8], z, wz); }
    st.put(SLOT + 0, wx);
    st.put(SLOT + 1, wy);
    st.put(SLOT + 2, wz);
  }
};

template <class T, class Store>
NWB_HD void load3(const Store& st, int slot, T* o) {
  o[0] = st.get(slot + 0);
  o[1] = st.get(slot + 1);
  o[2] = st.get(slot + 2);
}

// ---- exact squared distance segment (A + tD, t in [0,1]) -> box, in the box frame ---------------
// phi(t) = sum_k excess_k(t) * D_k is half the derivative of the convex piecewise-quadratic
// f(t) = sum_k excess_k(t)^2; it is non-decreasing and piecewise linear with break points at the
// (<= 6) face crossings.  Bracket its root between the break points, then interpolate linearly.
template <class T>
NWB_HD T box_excess(T x, T h) { return x - nwb_min(nwb_max(x, -h), h); }
template <class T>
NWB_HD T box_phi(const T* A, const T* D, const T* H, T t) {
  T s = T(0.f);
#pragma unroll
  for (int k = 0; k < 3; ++k) s = nwb_fma(box_excess(nwb_fma(t, D[k], A[k]), H[k]), D[k], s);
  return s;
}
template <class T>
NWB_HD T box_f(const T* A, const T* D, const T* H, T t) {
  T s = T(0.f);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T e = box_excess(nwb_fma(t, D[k], A[k]), H[k]);
    s = nwb_fma(e, e, s);
  }
  return s;
}
template <class T>
NWB_HD T seg_box_dist2(const T* A, const T* D, const T* H) {
  T lo = T(0.f), hi = T(1.f);
  const T p0 = box_phi(A, D, H, T(0.f)), p1 = box_phi(A, D, H, T(1.f));
  T plo = p0, phi = p1;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T inv = nwb_sel(D[k] != T(0.f), T(1.f) / D[k], T(0.f));
#pragma unroll
    for (int sg = 0; sg < 2; ++sg) {
      T face = sg ? H[k] : -H[k];
      T t = nwb_mul_sat(face - A[k], inv);
      T p = box_phi(A, D, H, t);
      auto below = nwb_and(p <= T(0.f), t >= lo);
      auto above = nwb_and(p >= T(0.f), t <= hi);
      lo = nwb_sel(below, t, lo); plo = nwb_sel(below, p, plo);
      hi = nwb_sel(above, t, hi); phi = nwb_sel(above, p, phi);
    }
  }
  T den = phi - plo;
  T t = nwb_sel(den > T(0.f), nwb_fma(hi - lo, nwb_sat(nwb_div_fast(-plo, den)), lo), lo);
  t = nwb_sel(p0 >= T(0.f), T(0.f), t);     // already moving away at the start
  t = nwb_sel(p1 <= T(0.f), T(1.f), t);     // still approaching at the end
  return box_f(A, D, H, t);
}

template <class T>
struct ValidityResult {
  typename mask_of<T>::type hit;      // some enabled pair / scene box in contact
  typename mask_of<T>::type stable;
};

// Pair-set selector: the default SRDF reading is compiled in (every slot, length and radius an
// immediate, loops fully unrolled); any other reading walks the runtime tables.
enum PairMode { PAIRS_TABLE = 0, PAIRS_STATIC_GROUP = 1, PAIRS_STATIC_ALL = 2 };

// Evaluate one configuration (or one lane-pair of configurations).  q = 21 joint values.
// MARGINS: also min separation / hull distance (one-lane scalars only).
// ENV_CULL: conservative bounding-sphere cull of the capsule-box tests under a warp vote.  It pays where the
// lanes of a warp are coherent (one lane of a planner kernel, consecutive way-points of an edge) and costs
// time on batches of unrelated configurations, whose warps almost never agree - the batched validity
// kernels instantiate it off.
template <class T, class Store, bool MARGINS, bool COLLISION_ONLY, int PAIRS = PAIRS_TABLE, bool ENV_CULL = false, bool BIG_SCENE = false>
NWB_HD ValidityResult<T> evaluate_config(const ValidityTables& tab, const T* q, Store& st, T* min_sep_out, T* hull_out,
                                         const SceneNode* staged_nodes = nullptr, int n_staged = 0) {
  using M = typename mask_of<T>::type;
  ValidityVisitor<T, Store, !COLLISION_ONLY> V(st);
  { NWB_MODEL_FK_WALK(T, XF<T>, q, V) }

  // ---- static stability: projected COM strictly inside hull(right foot U left foot) ------------
  M stable = nwb_mask<M>(true);
  T hull_margin = T(0.f);
  if constexpr (!COLLISION_ONLY) {
    const XF<T>& L = V.lsole;
    if constexpr (!MARGINS) {
      Wedge<T> w;
      w.init(T(tab.foot_right[0][0]) - V.comx, T(tab.foot_right[0][1]) - V.comy);
#pragma unroll
      for (int k = 1; k < nwbm::kFootN; ++k) w.add(T(tab.foot_right[k][0]) - V.comx, T(tab.foot_right[k][1]) - V.comy);
      // left-foot vertices relative to the COM: (L.p - com) + L.R * v
      const T ox = L.p[0] - V.comx, oy = L.p[1] - V.comy;
#pragma unroll
      for (int k = 0; k < nwbm::kFootN; ++k) {
        T x = T(tab.foot_left[k][0]), y = T(tab.foot_left[k][1]);
        w.add(nwb_fma(L.r[1], y, nwb_fma(L.r[0], x, ox)), nwb_fma(L.r[4], y, nwb_fma(L.r[3], x, oy)));
      }
      stable = w.surrounded;
    } else {
      T px[2 * nwbm::kFootN], py[2 * nwbm::kFootN];
#pragma unroll
      for (int k = 0; k < nwbm::kFootN; ++k) {
        px[k] = T(tab.foot_right[k][0]);
        py[k] = T(tab.foot_right[k][1]);
        T x = T(tab.foot_left[k][0]), y = T(tab.foot_left[k][1]);
        px[nwbm::kFootN + k] = nwb_fma(L.r[1], y, nwb_fma(L.r[0], x, L.p[0]));
        py[nwbm::kFootN + k] = nwb_fma(L.r[4], y, nwb_fma(L.r[3], x, L.p[1]));
      }
      hull_margin = hull_signed_distance<T, 2 * nwbm::kFootN>(px, py, V.comx, V.comy);
      stable = (hull_margin > T(0.f));
    }
  }

  // ---- self collision over the enabled pairs -----------------------------------------------------
  M hit = nwb_mask<M>(false);
  T min_sep = T(1e30f);
  auto account = [&](T d2, T rsum) {
    if constexpr (MARGINS) min_sep = nwb_min(min_sep, nwb_sqrt(d2) - rsum);
    else hit = nwb_or(hit, d2 < rsum * rsum);
  };
  // compiled-in pair list: the pair's (rA+rB)^2 is an immediate, so the distance routine returns d^2 - (rA+rB)^2
  // straight from its last multiply-add and contact = "the smallest of these is negative": one running minimum
  // (the compiler fuses two updates into one 3-input FMNMX3) instead of one compare per pair
  T worst = T(1e30f);
  auto account2 = [&](T d2_biased, T rsum, T rsum2) {
    if constexpr (MARGINS) min_sep = nwb_min(min_sep, nwb_sqrt(nwb_max(d2_biased + rsum2, T(0.f))) - rsum);
    else worst = nwb_min(worst, d2_biased);
  };
  (void)account2;
  if constexpr (PAIRS == PAIRS_TABLE) {
    int r = 0;
    for (; r < tab.n_runs_cc; ++r) {               // capsule - capsule
      const PairRun& run = tab.runs[r];
      T pa[3], qa[3], d1[3];
      load3(st, run.off_a, pa);
      load3(st, run.off_a + 3, qa);
      d1[0] = qa[0] - pa[0]; d1[1] = qa[1] - pa[1]; d1[2] = qa[2] - pa[2];
      const int cnt = run.count, first = run.first;
      const T a = T(run.a), inv_a = T(run.inv_a);
#pragma unroll 2
      for (int k = 0; k < cnt; ++k) {
        const PairEntry& pe = tab.entries[first + k];
        T pb[3], qb[3], d2[3];
        load3(st, pe.off_b, pb);
        load3(st, pe.off_b + 3, qb);
        d2[0] = qb[0] - pb[0]; d2[1] = qb[1] - pb[1]; d2[2] = qb[2] - pb[2];
        account(seg_seg_dist2(pa, d1, a, inv_a, pb, d2, T(pe.e), T(pe.inv_e)), T(pe.rsum));
      }
    }
    for (; r < tab.n_runs_cc + tab.n_runs_cs; ++r) {   // capsule - sphere
      const PairRun& run = tab.runs[r];
      T pa[3], qa[3], d1[3];
      load3(st, run.off_a, pa);
      load3(st, run.off_a + 3, qa);
      d1[0] = qa[0] - pa[0]; d1[1] = qa[1] - pa[1]; d1[2] = qa[2] - pa[2];
      const int cnt = run.count, first = run.first;
      const T inv_a = T(run.inv_a);
#pragma unroll 2
      for (int k = 0; k < cnt; ++k) {
        const PairEntry& pe = tab.entries[first + k];
        T pb[3];
        load3(st, pe.off_b, pb);
        account(seg_point_dist2(pa, d1, inv_a, pb), T(pe.rsum));
      }
    }
    for (; r < tab.n_runs_cc + tab.n_runs_cs + tab.n_runs_ss; ++r) {   // sphere - sphere
      const PairRun& run = tab.runs[r];
      T pa[3];
      load3(st, run.off_a, pa);
      const int cnt = run.count, first = run.first;
#pragma unroll 2
      for (int k = 0; k < cnt; ++k) {
        const PairEntry& pe = tab.entries[first + k];
        T pb[3];
        load3(st, pe.off_b, pb);
        account(point_point_dist2(pa, pb), T(pe.rsum));
      }
    }
  } else {
#define NWB_X_CC(SA, LA, ILA, SB, LB, ILB, RS, RS2)                                        \
    {                                                                                        \
      T pa[3], qa[3], d1[3], pb[3], qb[3], d2[3];                                            \
      load3(st, SA, pa); load3(st, SA + 3, qa); load3(st, SB, pb); load3(st, SB + 3, qb);    \
      d1[0] = qa[0] - pa[0]; d1[1] = qa[1] - pa[1]; d1[2] = qa[2] - pa[2];                   \
      d2[0] = qb[0] - pb[0]; d2[1] = qb[1] - pb[1]; d2[2] = qb[2] - pb[2];                   \
      account2(seg_seg_dist2(pa, d1, T(LA), T(ILA), pb, d2, T(LB), T(ILB), T(-(RS2))), T(RS), T(RS2)); \
    }
#define NWB_X_CS(SA, LA, ILA, SB, LB, ILB, RS, RS2)                                        \
    {                                                                                        \
      T pa[3], qa[3], d1[3], pb[3];                                                          \
      load3(st, SA, pa); load3(st, SA + 3, qa); load3(st, SB, pb);                           \
      d1[0] = qa[0] - pa[0]; d1[1] = qa[1] - pa[1]; d1[2] = qa[2] - pa[2];                   \
      account2(seg_point_dist2(pa, d1, T(ILA), pb, T(-(RS2))), T(RS), T(RS2));               \
    }
#define NWB_X_SS(SA, LA, ILA, SB, LB, ILB, RS, RS2)                                        \
    {                                                                                        \
      T pa[3], pb[3];                                                                        \
      load3(st, SA, pa); load3(st, SB, pb);                                                  \
      account2(point_point_dist2(pa, pb, T(-(RS2))), T(RS), T(RS2));                         \
    }
    if constexpr (PAIRS == PAIRS_STATIC_GROUP) {
      NWB_PAIRS_CC_GROUP(NWB_X_CC)
      NWB_PAIRS_CS_GROUP(NWB_X_CS)
      NWB_PAIRS_SS_GROUP(NWB_X_SS)
    } else {
      NWB_PAIRS_CC_ALL(NWB_X_CC)
      NWB_PAIRS_CS_ALL(NWB_X_CS)
      NWB_PAIRS_SS_ALL(NWB_X_SS)
    }
#undef NWB_X_CC
#undef NWB_X_CS
#undef NWB_X_SS
    if constexpr (!MARGINS) hit = nwb_or(hit, worst < T(0.f));
  }

  // ---- environment: robot proxies against the scene boxes ---------------------------------------
  // One box against every tested link: exact capsule-box / sphere-box distance.
  auto test_box = [&](const SceneBox& bx) __attribute__((always_inline)) {
    const T H[3] = {T(bx.h[0]), T(bx.h[1]), T(bx.h[2])};
    auto env_link = [&](int off, bool is_sphere, T radius, float half_len) {
      T pa[3], A[3];
      load3(st, off, pa);
      if constexpr (!MARGINS && ENV_CULL) {
        // Conservative cull: if the capsule's bounding sphere (mid point, half length + radius) is clear of the
        // box for EVERY lane of the warp, the exact segment-box distance cannot report contact - skip it.
        // Never changes a verdict; off (tab.env_cull = 0) for the full-evaluation roofline runs.
        if (!is_sphere && tab.env_cull) {
          T qm[3];
          load3(st, off + 3, qm);
          T mx = nwb_fma(T(0.5f), qm[0] - pa[0], pa[0]) - T(bx.c[0]);
          T my = nwb_fma(T(0.5f), qm[1] - pa[1], pa[1]) - T(bx.c[1]);
          T mz = nwb_fma(T(0.5f), qm[2] - pa[2], pa[2]) - T(bx.c[2]);
          T c0 = nwb_fma(T(bx.r[6]), mz, nwb_fma(T(bx.r[3]), my, T(bx.r[0]) * mx));
          T c1 = nwb_fma(T(bx.r[7]), mz, nwb_fma(T(bx.r[4]), my, T(bx.r[1]) * mx));
          T c2 = nwb_fma(T(bx.r[8]), mz, nwb_fma(T(bx.r[5]), my, T(bx.r[2]) * mx));
          T e0 = box_excess(c0, H[0]), e1 = box_excess(c1, H[1]), e2 = box_excess(c2, H[2]);
          T dm2 = nwb_fma(e2, e2, nwb_fma(e1, e1, e0 * e0));
          const float rb = half_len + 1.0e-4f;               // slack for the fp32 evaluation of the exact test
          if (nwb_warp_all(dm2 > (T(rb) + radius) * (T(rb) + radius))) return;
        }
      }
      T rx = pa[0] - T(bx.c[0]), ry = pa[1] - T(bx.c[1]), rz = pa[2] - T(bx.c[2]);
      A[0] = nwb_fma(T(bx.r[6]), rz, nwb_fma(T(bx.r[3]), ry, T(bx.r[0]) * rx));      // R^T (a - c)
      A[1] = nwb_fma(T(bx.r[7]), rz, nwb_fma(T(bx.r[4]), ry, T(bx.r[1]) * rx));
      A[2] = nwb_fma(T(bx.r[8]), rz, nwb_fma(T(bx.r[5]), ry, T(bx.r[2]) * rx));
      T d2;
      if (is_sphere) {                                                 // table / compile-time: warp-uniform
        T ex = box_excess(A[0], H[0]), ey = box_excess(A[1], H[1]), ez = box_excess(A[2], H[2]);
        d2 = nwb_fma(ez, ez, nwb_fma(ey, ey, ex * ex));
      } else {
        T qa[3], D[3];
        load3(st, off + 3, qa);
        T dx = qa[0] - pa[0], dy = qa[1] - pa[1], dz = qa[2] - pa[2];
        D[0] = nwb_fma(T(bx.r[6]), dz, nwb_fma(T(bx.r[3]), dy, T(bx.r[0]) * dx));
        D[1] = nwb_fma(T(bx.r[7]), dz, nwb_fma(T(bx.r[4]), dy, T(bx.r[1]) * dx));
        D[2] = nwb_fma(T(bx.r[8]), dz, nwb_fma(T(bx.r[5]), dy, T(bx.r[2]) * dx));
        d2 = seg_box_dist2(A, D, H);
      }
      account(d2, radius);
    };
    if constexpr (PAIRS == PAIRS_TABLE) {
      for (int l = 0; l < tab.n_env_links; ++l) {
        const EnvLink& el = tab.env_links[l];
        env_link(el.off, el.is_sphere != 0, T(el.radius), el.half_len);
      }
    } else {
#define NWB_X_ENV(SLOT, SPHERE, R, HALF) env_link(SLOT, SPHERE, T(R), HALF);
      if constexpr (PAIRS == PAIRS_STATIC_GROUP) { NWB_ENV_LINKS_GROUP(NWB_X_ENV) }
      else { NWB_ENV_LINKS_ALL(NWB_X_ENV) }
#undef NWB_X_ENV
    }
  };
  // inline boxes of the kernel parameters (scenes of <= kMaxBoxes boxes)
  for (int b = 0; b < tab.n_boxes; ++b) test_box(tab.boxes[b]);
  // Larger scenes (BIG_SCENE instantiations, float only): boxes in global memory under a flattened bounding-volume
  // hierarchy.  The walk visits the leaves whose bounds the robot's own bounding box (all tested proxies, radii
  // included) reaches and runs the SAME exact test on them, so the hierarchy never changes a verdict - it only skips
  // boxes that are clear of the whole robot.  MARGINS: a subtree is also entered while it could still lower the minimum
  // separation (branch and bound), so the reported minimum stays exact.
  if constexpr (BIG_SCENE && std::is_same<T, float>::value) {
    if (tab.scene.n_nodes > 0) {
      float rlo[3] = {1e30f, 1e30f, 1e30f}, rhi[3] = {-1e30f, -1e30f, -1e30f};
      auto grow = [&](int off, bool is_sphere, float radius) {
        const float rr = radius + 1.0e-4f;                              // slack for the fp32 evaluation of the exact test
        float p[3];
        load3(st, off, p);
#pragma unroll
        for (int k = 0; k < 3; ++k) { rlo[k] = fminf(rlo[k], p[k] - rr); rhi[k] = fmaxf(rhi[k], p[k] + rr); }
        if (!is_sphere) {
          load3(st, off + 3, p);
#pragma unroll
          for (int k = 0; k < 3; ++k) { rlo[k] = fminf(rlo[k], p[k] - rr); rhi[k] = fmaxf(rhi[k], p[k] + rr); }
        }
      };
      if constexpr (PAIRS == PAIRS_TABLE) {
        for (int l = 0; l < tab.n_env_links; ++l) grow(tab.env_links[l].off, tab.env_links[l].is_sphere != 0, tab.env_links[l].radius);
      } else {
#define NWB_X_GROW(SLOT, SPHERE, R, HALF) grow(SLOT, SPHERE, R);
        if constexpr (PAIRS == PAIRS_STATIC_GROUP) { NWB_ENV_LINKS_GROUP(NWB_X_GROW) }
        else { NWB_ENV_LINKS_ALL(NWB_X_GROW) }
#undef NWB_X_GROW
      }
      int node = 0;
      while (node < tab.scene.n_nodes) {
        if constexpr (!MARGINS) {
          if (hit) break;                                               // verdict settled: the rest of the scene cannot change it
        }
        // the first n_staged nodes of the hierarchy sit in shared memory (bulk-copied once per CTA by the batched kernel)
        const SceneNode nd = node < n_staged ? staged_nodes[node] : tab.scene.nodes[node];
        float g2 = 0.f;                                                 // squared gap between the two bounding boxes
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float g = fmaxf(0.f, fmaxf(nd.lo[k] - rhi[k], rlo[k] - nd.hi[k]));
          g2 = fmaf(g, g, g2);
        }
        bool visit = g2 <= 0.f;
        if constexpr (MARGINS) visit = visit || (min_sep > 0.f && g2 < min_sep * min_sep);
        if (!visit) { node = nd.skip; continue; }
        ++node;
        if (nd.box >= 0) {
          const SceneBox bx = tab.scene.boxes[nd.box];
          test_box(bx);
        }
      }
    }
  }
  if constexpr (MARGINS) {
    hit = (min_sep < T(0.f));
    if (min_sep_out) *min_sep_out = min_sep;
    if (hull_out) *hull_out = hull_margin;
  }
  ValidityResult<T> res;
  res.hit = hit;
  res.stable = stable;
  return res;
}

}  // namespace nwb
