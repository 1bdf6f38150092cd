// TEST INFRASTRUCTURE ONLY.  C shim around the reference's generated right-arm solver, compiled from
// the reference source WHERE IT LIES (-I/root/reference/rrt_connect_planner/src); nothing of that
// file is copied here.  Built only when /root/reference exists, output oracle/_ref/libikfast_ref.so.
// Used by tests/ to pin oracle/arm_ik_restated.h (our own closed-form restatement) against the
// solver the reference links (ikfast_TranslationDirection5D.cpp:341-1517 through the entry :1521).
#define IKFAST_NO_MAIN
#define IKFAST_NAMESPACE nwb_ikf_ref
#include "ikfast_TranslationDirection5D.cpp"

extern "C" {

// pos[3], dir[3] in the solver's torso frame -> up to `cap` solutions of 5 joints; returns the count
// (0 when the solver reports failure).  Order = the solver's own order.
__attribute__((visibility("default"))) int ikf_ref_solve(const double* pos, const double* dir, double* out, int cap) {
  double rot[9] = {dir[0], dir[1], dir[2], 0, 0, 0, 0, 0, 0};
  double pfree[5] = {5.0, 1.0, 1.0, 1.0, 1.0};   // as NCK:396-401 passes them (no free joints)
  std::vector<nwb_ikf_ref::IKSolution> sols;
  if (!nwb_ikf_ref::ik(pos, rot, pfree, sols)) return 0;
  int n = 0;
  for (size_t s = 0; s < sols.size() && n < cap; ++s, ++n) {
    double q[5];
    sols[s].GetSolution(q, nullptr);
    for (int j = 0; j < 5; ++j) out[n * 5 + j] = q[j];
  }
  return n;
}

__attribute__((visibility("default"))) void ikf_ref_fk(const double* q, double* pos, double* dir) {
  double rot[9];
  nwb_ikf_ref::fk(q, pos, rot);
  dir[0] = rot[0]; dir[1] = rot[1]; dir[2] = rot[2];
}

}
