// count_flops.cpp - freezes the ALGORITHMIC work per configuration of the validity kernel.
//
// Instantiates the exact device algorithm (csrc/validity_core.cuh) on the host with a scalar type
// that counts floating-point operations, over a sample of configurations, and prints JSON.
// Convention: add/sub/mul/div/sqrt/min/max/saturate/compare = 1 flop,
// sincos = 2 flop; an FMA therefore counts 2 (mul + add).  Selects, negation, copies and integer /
// boolean work are free.  Build + run: tools/run_count_flops.sh  (writes profiles/roofline.json).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

namespace nwb {
struct Counted {
  double v;
  static long long flops;
  Counted() : v(0) {}
  explicit Counted(double x) : v(x) {}
  explicit Counted(float x) : v(x) {}
  explicit Counted(int x) : v(x) {}
};
long long Counted::flops = 0;
inline Counted operator+(Counted a, Counted b) { ++Counted::flops; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { ++Counted::flops; return Counted(a.v - b.v); }
inline Counted operator*(Counted a, Counted b) { ++Counted::flops; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { ++Counted::flops; return Counted(a.v / b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }
inline Counted& operator+=(Counted& a, Counted b) { ++Counted::flops; a.v += b.v; return a; }
inline Counted& operator-=(Counted& a, Counted b) { ++Counted::flops; a.v -= b.v; return a; }
inline bool operator<(Counted a, Counted b) { ++Counted::flops; return a.v < b.v; }
inline bool operator>(Counted a, Counted b) { ++Counted::flops; return a.v > b.v; }
inline bool operator<=(Counted a, Counted b) { ++Counted::flops; return a.v <= b.v; }
inline bool operator>=(Counted a, Counted b) { ++Counted::flops; return a.v >= b.v; }
inline bool operator==(Counted a, Counted b) { ++Counted::flops; return a.v == b.v; }
inline bool operator!=(Counted a, Counted b) { ++Counted::flops; return a.v != b.v; }
inline void nwb_sincos(Counted a, Counted& s, Counted& c) { Counted::flops += 2; s = Counted(std::sin(a.v)); c = Counted(std::cos(a.v)); }
inline Counted nwb_sat(Counted x) { ++Counted::flops; return Counted(x.v < 0 ? 0.0 : (x.v > 1 ? 1.0 : x.v)); }
inline Counted nwb_sqrt(Counted x) { ++Counted::flops; return Counted(std::sqrt(x.v)); }
inline Counted nwb_min(Counted a, Counted b) { ++Counted::flops; return Counted(std::fmin(a.v, b.v)); }
inline Counted nwb_max(Counted a, Counted b) { ++Counted::flops; return Counted(std::fmax(a.v, b.v)); }
inline Counted nwb_fma(Counted a, Counted b, Counted c) { Counted::flops += 2; return Counted(a.v * b.v + c.v); }
inline Counted nwb_mul_sat(Counted a, Counted b) { Counted::flops += 2; double x = a.v * b.v; return Counted(x < 0 ? 0.0 : (x > 1 ? 1.0 : x)); }
inline Counted nwb_sel(bool c, Counted a, Counted b) { return c ? a : b; }
inline Counted nwb_div_fast(Counted a, Counted b) { ++Counted::flops; return Counted(a.v / b.v); }
}  // namespace nwb

#include "../nao_wholebody_planning_b200/csrc/nwb_tables.h"
#include "../nao_wholebody_planning_b200/csrc/validity_core.cuh"

using namespace nwb;

struct ArrayStore {
  Counted pts[nwbm::kPointFloats];
  void put(int slot, Counted v) { pts[slot] = v; }
  Counted get(int slot) const { return pts[slot]; }
};

static std::vector<SceneBox> shelf_scene() {   // S1 of SURVEY.md Appendix D
  const double sx = 0.33, sy = -0.06, sz = 0.02;
  double b[7][6] = {{sx, sy, sz, 0.27, 0.31, 0.04},
                    {sx + 0.0175, sy + 0.132, sz + 0.28, 0.235, 0.014, 0.52},
                    {sx + 0.0175, sy - 0.132, sz + 0.28, 0.235, 0.014, 0.52},
                    {sx, sy, sz + 0.5485, 0.27, 0.31, 0.017},
                    {sx + 0.025, sy, sz + 0.147, 0.22, 0.25, 0.014},
                    {sx + 0.025, sy, sz + 0.2775, 0.22, 0.25, 0.014},
                    {sx + 0.025, sy, sz + 0.408, 0.22, 0.25, 0.014}};
  std::vector<SceneBox> v;
  for (auto& r : b) {
    SceneBox s{};
    for (int k = 0; k < 3; ++k) { s.c[k] = (float)r[k]; s.h[k] = (float)(r[3 + k] * 0.5); }
    s.r[0] = s.r[4] = s.r[8] = 1.f;
    v.push_back(s);
  }
  return v;
}

template <bool MARGINS, bool COLLISION_ONLY, int PAIRS = PAIRS_TABLE>
static double count(const ValidityTables& tab, int samples) {
  std::mt19937_64 rng(20121001);
  std::uniform_real_distribution<double> U(0, 1);
  long long total = 0;
  for (int i = 0; i < samples; ++i) {
    Counted q[NWB_NJ];
    for (int j = 0; j < NWB_NJ; ++j) q[j] = Counted(nwbm::kJointMin[j] + (nwbm::kJointMax[j] - nwbm::kJointMin[j]) * U(rng));
    ArrayStore st;
    Counted ms, hm;
    Counted::flops = 0;
    evaluate_config<Counted, ArrayStore, MARGINS, COLLISION_ONLY, PAIRS>(tab, q, st, &ms, &hm);
    total += Counted::flops;
  }
  return (double)total / samples;
}

int main() {
  nwb_params p{};
  p.srdf_trim = 1; p.scale_origin = 1; p.scale_sp = 0.8;
  ValidityTables s0g, s0a, s1g;
  build_validity_tables(p, {}, true, s0g);
  build_validity_tables(p, {}, false, s0a);
  build_validity_tables(p, shelf_scene(), true, s1g);
  s0g.env_cull = s0a.env_cull = s1g.env_cull = 0;      // the roofline counts the FULL evaluation (no culling)
  const int N = 2000;
  int ncc = 0, ncs = 0, nss = 0;
  for (int r = 0; r < s0g.n_runs_cc; ++r) ncc += s0g.runs[r].count;
  for (int r = s0g.n_runs_cc; r < s0g.n_runs_cc + s0g.n_runs_cs; ++r) ncs += s0g.runs[r].count;
  for (int r = s0g.n_runs_cc + s0g.n_runs_cs; r < s0g.n_runs_cc + s0g.n_runs_cs + s0g.n_runs_ss; ++r) nss += s0g.runs[r].count;
  printf("{\n");
  printf(" \"convention\": \"add/sub/mul/div/sqrt/min/max/saturate/compare = 1 flop, sincos = 2, FMA = 2; selects free\",\n");
  printf(" \"source\": \"tools/count_flops.cpp: csrc/validity_core.cuh instantiated with a counting scalar, mean over %d uniform configs\",\n", N);
  printf(" \"pairs\": {\"capsule_capsule\": %d, \"capsule_sphere\": %d, \"sphere_sphere\": %d},\n", ncc, ncs, nss);
  // the table-driven form computes each capsule direction once per run = the minimal algorithm;
  // the compiled-in pair list is the same arithmetic after common-subexpression elimination
  printf(" \"validity_S0_flops_per_config\": %.1f,\n", count<false, false>(s0g, N));
  printf(" \"validity_S0_static_pairs_as_written\": %.1f,\n", count<false, false, PAIRS_STATIC_GROUP>(s0g, N));
  printf(" \"validity_S0_margins_flops_per_config\": %.1f,\n", count<true, false>(s0g, N));
  printf(" \"collision_only_S0_all_pairs_flops_per_config\": %.1f,\n", count<false, true>(s0a, N));
  printf(" \"validity_S1_shelf_flops_per_config\": %.1f,\n", count<false, false>(s1g, N));
  printf(" \"bytes_per_config\": {\"in_f64\": %d, \"in_f32\": %d, \"out\": 2}\n", NWB_NJ * 8, NWB_NJ * 4);
  printf("}\n");
  return 0;
}
