// nwb_math.cuh - scalar-generic geometry used by the CUDA kernels.
//
// Everything is a template over the scalar T and __host__ __device__, so the exact device
// algorithm can also be instantiated on the host with a flop-counting scalar (tools/count_flops.cu)
// to freeze the "algorithmic flops per configuration" the roofline is computed from.
#pragma once
#include <math.h>

#include "nao_model_gen.h"

#if defined(__CUDACC__)
#define NWB_HD __host__ __device__ __forceinline__
#else
#define NWB_HD inline
#endif

namespace nwb {

// ---- scalar helpers (overloaded for float / double; a counting type provides its own) --------
NWB_HD void nwb_sincos(float a, float& s, float& c) {
#if defined(__CUDA_ARCH__) && defined(NWB_FAST_SINCOS)
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
NWB_HD float nwb_sat(float x) {
#if defined(__CUDA_ARCH__)
  return __saturatef(x);
#else
  return x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
#endif
}
NWB_HD double nwb_sat(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }
NWB_HD float nwb_sqrt(float x) { return sqrtf(x); }
NWB_HD double nwb_sqrt(double x) { return sqrt(x); }
NWB_HD float nwb_min(float a, float b) { return fminf(a, b); }
NWB_HD double nwb_min(double a, double b) { return fmin(a, b); }
NWB_HD float nwb_max(float a, float b) { return fmaxf(a, b); }
NWB_HD double nwb_max(double a, double b) { return fmax(a, b); }
NWB_HD float nwb_abs(float a) { return fabsf(a); }
NWB_HD double nwb_abs(double a) { return fabs(a); }
template <class T>
NWB_HD T nwb_sel(bool c, T a, T b) { return c ? a : b; }
// a / b where 2 ulp are enough (interpolation parameters that get clamped to [0,1] anyway):
// one MUFU.RCP + one FMUL on the device instead of the IEEE division sequence.
NWB_HD float nwb_div_fast(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdividef(a, b);
#else
  return a / b;
#endif
}
NWB_HD double nwb_div_fast(double a, double b) { return a / b; }

// ---- rigid transform: rotation (row-major r[3*row+col]) + translation, in the r_sole frame ----
template <class T>
struct XF {
  T r[9];
  T p[3];
};
template <class T>
NWB_HD XF<T> nwb_xf_identity() {
  XF<T> f;
  f.r[0] = T(1); f.r[1] = T(0); f.r[2] = T(0);
  f.r[3] = T(0); f.r[4] = T(1); f.r[5] = T(0);
  f.r[6] = T(0); f.r[7] = T(0); f.r[8] = T(1);
  f.p[0] = T(0); f.p[1] = T(0); f.p[2] = T(0);
  return f;
}
// translate along a local axis: p += d * R[:,k]
template <class T> NWB_HD void nwb_xf_tx(XF<T>& f, T d) { f.p[0] += d * f.r[0]; f.p[1] += d * f.r[3]; f.p[2] += d * f.r[6]; }
template <class T> NWB_HD void nwb_xf_ty(XF<T>& f, T d) { f.p[0] += d * f.r[1]; f.p[1] += d * f.r[4]; f.p[2] += d * f.r[7]; }
template <class T> NWB_HD void nwb_xf_tz(XF<T>& f, T d) { f.p[0] += d * f.r[2]; f.p[1] += d * f.r[5]; f.p[2] += d * f.r[8]; }
// R <- R * RotX(s,c): columns 1,2 mix
template <class T>
NWB_HD void nwb_xf_rotx(XF<T>& f, T s, T c) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T a = f.r[3 * k + 1], b = f.r[3 * k + 2];
    f.r[3 * k + 1] = c * a + s * b;
    f.r[3 * k + 2] = c * b - s * a;
  }
}
// R <- R * RotY(s,c): columns 0,2 mix
template <class T>
NWB_HD void nwb_xf_roty(XF<T>& f, T s, T c) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T a = f.r[3 * k + 0], b = f.r[3 * k + 2];
    f.r[3 * k + 0] = c * a - s * b;
    f.r[3 * k + 2] = s * a + c * b;
  }
}
// R <- R * RotZ(s,c): columns 0,1 mix
template <class T>
NWB_HD void nwb_xf_rotz(XF<T>& f, T s, T c) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T a = f.r[3 * k + 0], b = f.r[3 * k + 1];
    f.r[3 * k + 0] = c * a + s * b;
    f.r[3 * k + 1] = c * b - s * a;
  }
}
// R <- R * Rodrigues(axis, s, c) for a unit axis
template <class T>
NWB_HD void nwb_xf_rotaxis(XF<T>& f, T ax, T ay, T az, T s, T c) {
  T v = T(1) - c;
  T m[9] = {ax * ax * v + c,      ax * ay * v - az * s, ax * az * v + ay * s,
            ay * ax * v + az * s, ay * ay * v + c,      ay * az * v - ax * s,
            az * ax * v - ay * s, az * ay * v + ax * s, az * az * v + c};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T a = f.r[3 * k + 0], b = f.r[3 * k + 1], d = f.r[3 * k + 2];
    f.r[3 * k + 0] = a * m[0] + b * m[3] + d * m[6];
    f.r[3 * k + 1] = a * m[1] + b * m[4] + d * m[7];
    f.r[3 * k + 2] = a * m[2] + b * m[5] + d * m[8];
  }
}
// world = p + R * (x,y,z) with compile-time-zero components skipped by the caller
template <class T>
NWB_HD void nwb_xf_point(const XF<T>& f, T x, T y, T z, T& ox, T& oy, T& oz) {
  ox = f.p[0] + f.r[0] * x + f.r[1] * y + f.r[2] * z;
  oy = f.p[1] + f.r[3] * x + f.r[4] * y + f.r[5] * z;
  oz = f.p[2] + f.r[6] * x + f.r[7] * y + f.r[8] * z;
}

// ---- closest squared distance between two segments, lengths known at compile time -------------
// Segment A = pa + s*d1, B = pb + t*d2.  a = |d1|^2, e = |d2|^2 (link constants), inv_a = 1/a,
// inv_e = 1/e (0 for a degenerate segment = sphere centre).
// Branch-free: unconstrained optimum for s (clamped), best t for that s (clamped), then the best s
// for that t (clamped).  The last step reproduces both end-point cases of Ericson's
// ClosestPtSegmentSegment and leaves an interior optimum unchanged, without any select.
template <class T>
NWB_HD T seg_seg_dist2(const T* pa, const T* d1, T a, T inv_a, const T* pb, const T* d2, T e, T inv_e) {
  T rx = pa[0] - pb[0], ry = pa[1] - pb[1], rz = pa[2] - pb[2];
  T f = d2[0] * rx + d2[1] * ry + d2[2] * rz;
  T c = d1[0] * rx + d1[1] * ry + d1[2] * rz;
  T b = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2];
  T denom = a * e - b * b;
  T s = nwb_sel(denom > T(1e-12), nwb_sat(nwb_div_fast(b * f - c * e, denom)), T(0));
  T t = nwb_sat((b * s + f) * inv_e);
  s = nwb_sat((b * t - c) * inv_a);
  T vx = rx + d1[0] * s - d2[0] * t;
  T vy = ry + d1[1] * s - d2[1] * t;
  T vz = rz + d1[2] * s - d2[2] * t;
  return vx * vx + vy * vy + vz * vz;
}
// segment A vs point pb
template <class T>
NWB_HD T seg_point_dist2(const T* pa, const T* d1, T inv_a, const T* pb) {
  T rx = pa[0] - pb[0], ry = pa[1] - pb[1], rz = pa[2] - pb[2];
  T c = d1[0] * rx + d1[1] * ry + d1[2] * rz;
  T s = nwb_sat(-c * inv_a);
  T vx = rx + d1[0] * s, vy = ry + d1[1] * s, vz = rz + d1[2] * s;
  return vx * vx + vy * vy + vz * vz;
}
template <class T>
NWB_HD T point_point_dist2(const T* pa, const T* pb) {
  T rx = pa[0] - pb[0], ry = pa[1] - pb[1], rz = pa[2] - pb[2];
  return rx * rx + ry * ry + rz * rz;
}

// ---- "is the origin strictly inside the convex hull of the directions fed so far" -------------
// Incremental angular wedge: (R, L) bound the cone spanned so far (angle < pi).  A direction that
// is counter-clockwise of L and clockwise of R closes the cone past pi: the origin is surrounded.
// Equivalent to hull construction + point-in-hull (hrl_kinematics) away from the boundary.
template <class T>
struct Wedge {
  T lx, ly, rx, ry;
  bool surrounded;
  NWB_HD void init(T vx, T vy) { lx = rx = vx; ly = ry = vy; surrounded = false; }
  NWB_HD void add(T vx, T vy) {
    T cl = lx * vy - ly * vx;
    T cr = rx * vy - ry * vx;
    bool ccw_l = cl > T(0), cw_r = cr < T(0);
    surrounded = surrounded || (ccw_l && cw_r);
    bool ext_l = ccw_l && !cw_r, ext_r = cw_r && !ccw_l;
    lx = nwb_sel(ext_l, vx, lx); ly = nwb_sel(ext_l, vy, ly);
    rx = nwb_sel(ext_r, vx, rx); ry = nwb_sel(ext_r, vy, ry);
  }
};

// ---- exact signed distance of point c to the hull of N points (gift wrapping, register-only) --
// > 0 inside (distance to the nearest edge line), < 0 outside (-distance to the hull).
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
  T inside = T(1e30), dmin2 = T(1e30);
  bool in = true;
  for (int step = 0; step < N; ++step) {
    T bx = ux, by = uy, bd = T(0);   // best next vertex so far, its squared distance from (ux,uy)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      T ex = bx - ux, ey = by - uy, qx = px[i] - ux, qy = py[i] - uy;
      T cr = ex * qy - ey * qx;
      T qd = qx * qx + qy * qy;
      bool take = (cr < T(0)) || (cr == T(0) && qd > bd);
      bx = nwb_sel(take, px[i], bx); by = nwb_sel(take, py[i], by); bd = nwb_sel(take, qd, bd);
    }
    // hull edge u -> b (counter-clockwise): signed line distance and segment distance of c
    T ex = bx - ux, ey = by - uy;
    T len2 = ex * ex + ey * ey;
    if (len2 > T(0)) {
      T wx = cx - ux, wy = cy - uy;
      T t = (ex * wy - ey * wx) / nwb_sqrt(len2);
      in = in && (t > T(0));
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
