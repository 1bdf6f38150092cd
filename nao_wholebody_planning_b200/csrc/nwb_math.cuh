// nwb_math.cuh - scalar-generic geometry used by the CUDA kernels.
//
// Everything is a template over a "scalar" T and __host__ __device__, so the exact device algorithm
// is instantiated three ways:
//   float    one configuration per thread
//   F2       two configurations per thread on Blackwell's packed FP32 pipe instructions
//            (fma.rn.f32x2 / add.f32x2 / mul.f32x2 -> SASS FFMA2/FADD2/FMUL2): one dispatch slot per
//            two lanes of work, which frees issue bandwidth for the compare/select instructions
//   Counted  a flop-counting host scalar (tools/count_flops.cpp) that freezes the "algorithmic
//            flops per configuration" the roofline is computed from.
// A scalar provides: T(float), + - * /, unary -, comparisons returning mask_of<T>::type, and the
// nwb_* helpers below (fma, saturate, select, mask logic).
#pragma once
#include <math.h>

#include "nao_model_gen.h"

#if defined(__CUDACC__)
#define NWB_HD __host__ __device__ __forceinline__
#else
#define NWB_HD inline
#endif

namespace nwb {

template <class T>
struct mask_of {
  using type = bool;
};

// ---- float / double ----------------------------------------------------------------------------
NWB_HD void nwb_sincos(float a, float& s, float& c) {
#if defined(__CUDA_ARCH__) && defined(NWB_FAST_SINCOS)
  // MUFU.SIN / MUFU.COS: max abs error 2^-21.4 on [-pi, pi] (every NAO joint range lies inside);
  // measured effect on link poses < 2e-6 m (tests/test_validity_gpu.py::test_link_poses_and_com)
  s = __sinf(a); c = __cosf(a);
#elif defined(__CUDA_ARCH__)
  sincosf(a, &s, &c);
#else
  s = sinf(a); c = cosf(a);
#endif
}
NWB_HD void nwb_sincos(double a, double& s, double& c) {
#if defined(__CUDA_ARCH__)
  sincos(a, &s, &c);
#else
  s = sin(a); c = cos(a);
#endif
}
NWB_HD float nwb_fma(float a, float b, float c) { return fmaf(a, b, c); }
NWB_HD double nwb_fma(double a, double b, double c) { return fma(a, b, c); }
NWB_HD float nwb_sat(float x) {
#if defined(__CUDA_ARCH__)
  return __saturatef(x);
#else
  return x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
#endif
}
NWB_HD double nwb_sat(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }
NWB_HD float nwb_mul_sat(float a, float b) { return nwb_sat(a * b); }      // FMUL.SAT on the device
NWB_HD double nwb_mul_sat(double a, double b) { return nwb_sat(a * b); }
NWB_HD float nwb_sqrt(float x) { return sqrtf(x); }
NWB_HD double nwb_sqrt(double x) { return sqrt(x); }
NWB_HD float nwb_min(float a, float b) { return fminf(a, b); }
NWB_HD double nwb_min(double a, double b) { return fmin(a, b); }
NWB_HD float nwb_max(float a, float b) { return fmaxf(a, b); }
NWB_HD double nwb_max(double a, double b) { return fmax(a, b); }
NWB_HD float nwb_sel(bool c, float a, float b) { return c ? a : b; }
NWB_HD double nwb_sel(bool c, double a, double b) { return c ? a : b; }
// a / b where 2 ulp are enough (interpolation parameters that get clamped to [0,1] anyway):
// one MUFU.RCP + one FMUL on the device instead of the IEEE division sequence.
NWB_HD float nwb_div_fast(float a, float b) {
#if defined(__CUDA_ARCH__)
  // bare MUFU.RCP + FMUL: __fdividef wraps the same two instructions in a range check for |b| > 2^126 (one FSETP and
  // two more FMULs per division); every denominator here is a squared length of at most a few m^2
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  return a * r;
#else
  return a / b;
#endif
}
NWB_HD double nwb_div_fast(double a, double b) { return a / b; }
// mask logic (bool for one-lane scalars)
NWB_HD bool nwb_and(bool a, bool b) { return a && b; }
NWB_HD bool nwb_or(bool a, bool b) { return a || b; }
NWB_HD bool nwb_not(bool a) { return !a; }
NWB_HD bool nwb_any(bool a) { return a; }

// ---- F2: two configurations per thread on the packed FP32 instructions --------------------------
#if defined(__CUDACC__)
struct M2 {
  bool x, y;
};
struct F2 {
  float2 v;
  F2() = default;
  __device__ __forceinline__ explicit F2(float s) : v(make_float2(s, s)) {}
  __device__ __forceinline__ F2(float a, float b) : v(make_float2(a, b)) {}
};
template <>
struct mask_of<F2> {
  using type = M2;
};
__device__ __forceinline__ F2 f2(float2 v) { F2 r; r.v = v; return r; }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { return f2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return f2(__fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y))); }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { return f2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ F2 operator/(F2 a, F2 b) { return F2(a.v.x / b.v.x, a.v.y / b.v.y); }
__device__ __forceinline__ F2 operator-(F2 a) { return F2(-a.v.x, -a.v.y); }
__device__ __forceinline__ F2& operator+=(F2& a, F2 b) { a = a + b; return a; }
__device__ __forceinline__ F2& operator-=(F2& a, F2 b) { a = a - b; return a; }
__device__ __forceinline__ M2 operator<(F2 a, F2 b) { return M2{a.v.x < b.v.x, a.v.y < b.v.y}; }
__device__ __forceinline__ M2 operator>(F2 a, F2 b) { return M2{a.v.x > b.v.x, a.v.y > b.v.y}; }
__device__ __forceinline__ M2 operator<=(F2 a, F2 b) { return M2{a.v.x <= b.v.x, a.v.y <= b.v.y}; }
__device__ __forceinline__ M2 operator>=(F2 a, F2 b) { return M2{a.v.x >= b.v.x, a.v.y >= b.v.y}; }
__device__ __forceinline__ M2 operator!=(F2 a, F2 b) { return M2{a.v.x != b.v.x, a.v.y != b.v.y}; }
__device__ __forceinline__ F2 nwb_fma(F2 a, F2 b, F2 c) { return f2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ void nwb_sincos(F2 a, F2& s, F2& c) {
  nwb_sincos(a.v.x, s.v.x, c.v.x);
  nwb_sincos(a.v.y, s.v.y, c.v.y);
}
__device__ __forceinline__ F2 nwb_sat(F2 a) { return F2(__saturatef(a.v.x), __saturatef(a.v.y)); }
__device__ __forceinline__ F2 nwb_mul_sat(F2 a, F2 b) { return F2(__saturatef(a.v.x * b.v.x), __saturatef(a.v.y * b.v.y)); }
__device__ __forceinline__ F2 nwb_sqrt(F2 a) { return F2(sqrtf(a.v.x), sqrtf(a.v.y)); }
__device__ __forceinline__ F2 nwb_min(F2 a, F2 b) { return F2(fminf(a.v.x, b.v.x), fminf(a.v.y, b.v.y)); }
__device__ __forceinline__ F2 nwb_max(F2 a, F2 b) { return F2(fmaxf(a.v.x, b.v.x), fmaxf(a.v.y, b.v.y)); }
__device__ __forceinline__ F2 nwb_sel(M2 m, F2 a, F2 b) { return F2(m.x ? a.v.x : b.v.x, m.y ? a.v.y : b.v.y); }
__device__ __forceinline__ F2 nwb_div_fast(F2 a, F2 b) { return F2(__fdividef(a.v.x, b.v.x), __fdividef(a.v.y, b.v.y)); }
__device__ __forceinline__ M2 nwb_and(M2 a, M2 b) { return M2{a.x && b.x, a.y && b.y}; }
__device__ __forceinline__ M2 nwb_or(M2 a, M2 b) { return M2{a.x || b.x, a.y || b.y}; }
__device__ __forceinline__ M2 nwb_not(M2 a) { return M2{!a.x, !a.y}; }
__device__ __forceinline__ bool nwb_any(M2 a) { return a.x || a.y; }
#endif

template <class M>
NWB_HD M nwb_mask(bool v);
template <>
NWB_HD bool nwb_mask<bool>(bool v) { return v; }
// "every active lane of the warp agrees": a real vote for the scalar device path, false elsewhere (host
// instantiations such as the flop counter, packed lanes), so a cull guarded by it is simply not taken there
template <class M>
NWB_HD bool nwb_warp_all(M) { return false; }
NWB_HD bool nwb_warp_all(bool v) {
#ifdef __CUDA_ARCH__
  return __all_sync(__activemask(), v);
#else
  (void)v;
  return false;
#endif
}
#if defined(__CUDACC__)
template <>
__device__ __forceinline__ M2 nwb_mask<M2>(bool v) { return M2{v, v}; }
#endif

// ---- rigid transform: rotation (row-major r[3*row+col]) + translation, in the r_sole frame ----
template <class T>
struct XF {
  T r[9];
  T p[3];
};
template <class T>
NWB_HD XF<T> nwb_xf_identity() {
  XF<T> f;
  f.r[0] = T(1.f); f.r[1] = T(0.f); f.r[2] = T(0.f);
  f.r[3] = T(0.f); f.r[4] = T(1.f); f.r[5] = T(0.f);
  f.r[6] = T(0.f); f.r[7] = T(0.f); f.r[8] = T(1.f);
  f.p[0] = T(0.f); f.p[1] = T(0.f); f.p[2] = T(0.f);
  return f;
}
// translate along a local axis: p += d * R[:,k]
template <class T> NWB_HD void nwb_xf_tx(XF<T>& f, T d) { f.p[0] = nwb_fma(d, f.r[0], f.p[0]); f.p[1] = nwb_fma(d, f.r[3], f.p[1]); f.p[2] = nwb_fma(d, f.r[6], f.p[2]); }
template <class T> NWB_HD void nwb_xf_ty(XF<T>& f, T d) { f.p[0] = nwb_fma(d, f.r[1], f.p[0]); f.p[1] = nwb_fma(d, f.r[4], f.p[1]); f.p[2] = nwb_fma(d, f.r[7], f.p[2]); }
template <class T> NWB_HD void nwb_xf_tz(XF<T>& f, T d) { f.p[0] = nwb_fma(d, f.r[2], f.p[0]); f.p[1] = nwb_fma(d, f.r[5], f.p[1]); f.p[2] = nwb_fma(d, f.r[8], f.p[2]); }
// R <- R * RotX(s,c): columns 1,2 mix
template <class T>
NWB_HD void nwb_xf_rotx(XF<T>& f, T s, T c) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T a = f.r[3 * k + 1], b = f.r[3 * k + 2];
    f.r[3 * k + 1] = nwb_fma(c, a, s * b);
    f.r[3 * k + 2] = nwb_fma(c, b, -(s * a));
  }
}
// R <- R * RotY(s,c): columns 0,2 mix
template <class T>
NWB_HD void nwb_xf_roty(XF<T>& f, T s, T c) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T a = f.r[3 * k + 0], b = f.r[3 * k + 2];
    f.r[3 * k + 0] = nwb_fma(c, a, -(s * b));
    f.r[3 * k + 2] = nwb_fma(s, a, c * b);
  }
}
// R <- R * RotZ(s,c): columns 0,1 mix
template <class T>
NWB_HD void nwb_xf_rotz(XF<T>& f, T s, T c) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T a = f.r[3 * k + 0], b = f.r[3 * k + 1];
    f.r[3 * k + 0] = nwb_fma(c, a, s * b);
    f.r[3 * k + 1] = nwb_fma(c, b, -(s * a));
  }
}
// R <- R * Rodrigues(axis, s, c) for a unit axis
template <class T>
NWB_HD void nwb_xf_rotaxis(XF<T>& f, T ax, T ay, T az, T s, T c) {
  T v = T(1.f) - c;
  T m[9] = {nwb_fma(ax * ax, v, c),       nwb_fma(ax * ay, v, -(az * s)), nwb_fma(ax * az, v, ay * s),
            nwb_fma(ay * ax, v, az * s),  nwb_fma(ay * ay, v, c),         nwb_fma(ay * az, v, -(ax * s)),
            nwb_fma(az * ax, v, -(ay * s)), nwb_fma(az * ay, v, ax * s),  nwb_fma(az * az, v, c)};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T a = f.r[3 * k + 0], b = f.r[3 * k + 1], d = f.r[3 * k + 2];
    f.r[3 * k + 0] = nwb_fma(a, m[0], nwb_fma(b, m[3], d * m[6]));
    f.r[3 * k + 1] = nwb_fma(a, m[1], nwb_fma(b, m[4], d * m[7]));
    f.r[3 * k + 2] = nwb_fma(a, m[2], nwb_fma(b, m[5], d * m[8]));
  }
}

template <class T>
NWB_HD T dot3(const T* a, T bx, T by, T bz) { return nwb_fma(a[2], bz, nwb_fma(a[1], by, a[0] * bx)); }

// ---- closest squared distance between two segments, lengths known at compile time -------------
// Segment A = pa + s*d1, B = pb + t*d2.  a = |d1|^2, e = |d2|^2 (link constants), inv_a = 1/a,
// inv_e = 1/e (0 for a degenerate segment = sphere centre).
// Branch- and select-free: unconstrained optimum for s (clamped), best t for that s (clamped), then
// the best s for that t (clamped).  The last step reproduces both end-point cases of Ericson's
// ClosestPtSegmentSegment and leaves an interior optimum unchanged.  The tiny bias on the
// denominator keeps parallel segments finite; whatever s they get is repaired by the two
// refinement steps (any s is optimal along parallel lines).
// `bias` is added to the result inside the last multiply-add chain: with bias = -(rA + rB)^2 the sign of the result is
// the contact test and no separate compare (an ALU-pipe instruction, two dispatch cycles on sm_100) is needed per pair.
template <class T>
NWB_HD T seg_seg_dist2(const T* pa, const T* d1, T a, T inv_a, const T* pb, const T* d2, T e, T inv_e, T bias) {
  T rx = pa[0] - pb[0], ry = pa[1] - pb[1], rz = pa[2] - pb[2];
  T f = dot3(d2, rx, ry, rz);
  T c = dot3(d1, rx, ry, rz);
  T b = dot3(d1, d2[0], d2[1], d2[2]);
  T denom = nwb_fma(-b, b, nwb_fma(a, e, T(1e-18f)));
  T s = nwb_sat(nwb_div_fast(nwb_fma(b, f, -(c * e)), denom));
  T t = nwb_mul_sat(nwb_fma(b, s, f), inv_e);
  s = nwb_mul_sat(nwb_fma(b, t, -c), inv_a);
  T vx = nwb_fma(-d2[0], t, nwb_fma(d1[0], s, rx));
  T vy = nwb_fma(-d2[1], t, nwb_fma(d1[1], s, ry));
  T vz = nwb_fma(-d2[2], t, nwb_fma(d1[2], s, rz));
  return nwb_fma(vz, vz, nwb_fma(vy, vy, nwb_fma(vx, vx, bias)));
}
// segment A vs point pb
template <class T>
NWB_HD T seg_point_dist2(const T* pa, const T* d1, T inv_a, const T* pb, T bias) {
  T rx = pa[0] - pb[0], ry = pa[1] - pb[1], rz = pa[2] - pb[2];
  T c = dot3(d1, rx, ry, rz);
  T s = nwb_mul_sat(-c, inv_a);
  T vx = nwb_fma(d1[0], s, rx), vy = nwb_fma(d1[1], s, ry), vz = nwb_fma(d1[2], s, rz);
  return nwb_fma(vz, vz, nwb_fma(vy, vy, nwb_fma(vx, vx, bias)));
}
template <class T>
NWB_HD T point_point_dist2(const T* pa, const T* pb, T bias) {
  T rx = pa[0] - pb[0], ry = pa[1] - pb[1], rz = pa[2] - pb[2];
  return nwb_fma(rz, rz, nwb_fma(ry, ry, nwb_fma(rx, rx, bias)));
}

// unbiased forms (plain squared distances: table-driven pair lists, margins, planner kernels)
template <class T>
NWB_HD T seg_seg_dist2(const T* pa, const T* d1, T a, T inv_a, const T* pb, const T* d2, T e, T inv_e) {
  T rx = pa[0] - pb[0], ry = pa[1] - pb[1], rz = pa[2] - pb[2];
  T f = dot3(d2, rx, ry, rz);
  T c = dot3(d1, rx, ry, rz);
  T b = dot3(d1, d2[0], d2[1], d2[2]);
  T denom = nwb_fma(-b, b, nwb_fma(a, e, T(1e-18f)));
  T s = nwb_sat(nwb_div_fast(nwb_fma(b, f, -(c * e)), denom));
  T t = nwb_mul_sat(nwb_fma(b, s, f), inv_e);
  s = nwb_mul_sat(nwb_fma(b, t, -c), inv_a);
  T vx = nwb_fma(-d2[0], t, nwb_fma(d1[0], s, rx));
  T vy = nwb_fma(-d2[1], t, nwb_fma(d1[1], s, ry));
  T vz = nwb_fma(-d2[2], t, nwb_fma(d1[2], s, rz));
  return nwb_fma(vz, vz, nwb_fma(vy, vy, vx * vx));
}
template <class T>
NWB_HD T seg_point_dist2(const T* pa, const T* d1, T inv_a, const T* pb) {
  T rx = pa[0] - pb[0], ry = pa[1] - pb[1], rz = pa[2] - pb[2];
  T c = dot3(d1, rx, ry, rz);
  T s = nwb_mul_sat(-c, inv_a);
  T vx = nwb_fma(d1[0], s, rx), vy = nwb_fma(d1[1], s, ry), vz = nwb_fma(d1[2], s, rz);
  return nwb_fma(vz, vz, nwb_fma(vy, vy, vx * vx));
}
template <class T>
NWB_HD T point_point_dist2(const T* pa, const T* pb) {
  T rx = pa[0] - pb[0], ry = pa[1] - pb[1], rz = pa[2] - pb[2];
  return nwb_fma(rz, rz, nwb_fma(ry, ry, rx * rx));
}

// ---- "is the origin strictly inside the convex hull of the directions fed so far" -------------
// Incremental angular wedge: (R, L) bound the cone spanned so far (angle < pi).  A direction that
// is counter-clockwise of L and clockwise of R closes the cone past pi: the origin is surrounded.
// Equivalent to hull construction + point-in-hull (hrl_kinematics) away from the boundary.
template <class T>
struct Wedge {
  using M = typename mask_of<T>::type;
  T lx, ly, rx, ry;
  M surrounded;
  NWB_HD void init(T vx, T vy) { lx = rx = vx; ly = ry = vy; surrounded = nwb_mask<M>(false); }
  NWB_HD void add(T vx, T vy) {
    T cl = nwb_fma(lx, vy, -(ly * vx));
    T cr = nwb_fma(rx, vy, -(ry * vx));
    M ccw_l = cl > T(0.f), cw_r = cr < T(0.f);
    surrounded = nwb_or(surrounded, nwb_and(ccw_l, cw_r));
    M ext_l = nwb_and(ccw_l, nwb_not(cw_r)), ext_r = nwb_and(cw_r, nwb_not(ccw_l));
    lx = nwb_sel(ext_l, vx, lx); ly = nwb_sel(ext_l, vy, ly);
    rx = nwb_sel(ext_r, vx, rx); ry = nwb_sel(ext_r, vy, ry);
  }
};

// ---- exact signed distance of point c to the hull of N points (gift wrapping, register-only) --
// > 0 inside (distance to the nearest edge line), < 0 outside (-distance to the hull).
// One-lane scalars only (data-dependent trip count).
template <class T, int N>
NWB_HD T hull_signed_distance(const T* px, const T* py, T cx, T cy) {
  // start: lexicographically smallest point
  T sx = px[0], sy = py[0];
#pragma unroll
  for (int i = 1; i < N; ++i) {
    bool less = (px[i] < sx) || (px[i] == sx && py[i] < sy);
    sx = nwb_sel(less, px[i], sx); sy = nwb_sel(less, py[i], sy);
  }
  T ux = sx, uy = sy;            // current hull vertex
  T inside = T(1e30f), dmin2 = T(1e30f);
  bool in = true;
  for (int step = 0; step < N; ++step) {
    T bx = ux, by = uy, bd = T(0.f);   // best next vertex so far, its squared distance from (ux,uy)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      T ex = bx - ux, ey = by - uy, qx = px[i] - ux, qy = py[i] - uy;
      T cr = ex * qy - ey * qx;
      T qd = qx * qx + qy * qy;
      bool take = (cr < T(0.f)) || (cr == T(0.f) && qd > bd);
      bx = nwb_sel(take, px[i], bx); by = nwb_sel(take, py[i], by); bd = nwb_sel(take, qd, bd);
    }
    // hull edge u -> b (counter-clockwise): signed line distance and segment distance of c
    T ex = bx - ux, ey = by - uy;
    T len2 = ex * ex + ey * ey;
    if (len2 > T(0.f)) {
      T wx = cx - ux, wy = cy - uy;
      T t = (ex * wy - ey * wx) / nwb_sqrt(len2);
      in = in && (t > T(0.f));
      inside = nwb_min(inside, t);
      T u = nwb_sat((wx * ex + wy * ey) / len2);
      T gx = wx - u * ex, gy = wy - u * ey;
      dmin2 = nwb_min(dmin2, gx * gx + gy * gy);
    }
    ux = bx; uy = by;
    if (ux == sx && uy == sy) break;
  }
  return in ? inside : -nwb_sqrt(dmin2);
}

}  // namespace nwb
