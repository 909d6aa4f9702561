// TEST INFRASTRUCTURE ONLY (oracle).  The RTR / RSD / NSD control flow of
// sagecal_b200/csrc/rtr_algo.h instantiated with the plain O(rows) CPU evaluators of
// dirac_oracle.c (orc_rtr_raw / orc_rtr_counts / orc_rtr_weights), so that
//   * the control flow is pinned against the compiled reference's rtr_solve_nocuda,
//     rtr_solve_nocuda_robust and nsd_solve_nocuda_robust without a GPU (tests/test_oracle_rtr.py);
//   * orc_sagefit can run solver_mode 4-6 (the harness registers itself with liboracle.so on load).
// The product never loads this library; its own evaluator is the device one (rtr.cu).
#include <vector>

#include "../sagecal_b200/csrc/rtr_algo.h"

extern "C" {
// dirac_oracle.h is C99 (double complex): the few prototypes needed here, with the problem opaque
double orc_rtr_raw(const void *P, int k, int t0, int ntiles, const double *y, const double *wt,
                   const double *x, const double *eta, double *vec);
void orc_rtr_counts(const void *P, int t0, int ntiles, double *cnt);
double orc_rtr_weights(const void *P, int k, int t0, int ntiles, const double *y, const double *x,
                       double nu, double *wt);
typedef void (*orc_rtr_solver_fn)(const void *P, int k, int t0, int ntiles, const double *y,
                                  int kind, double *x, int itmax_a, int itmax_b, double nulow,
                                  double nuhigh, double *robust_nu, double *info);
void orc_set_rtr_solver(orc_rtr_solver_fn fn);
}

namespace {
struct CpuEval {
  const void *P;
  int k, t0, ntiles, N;
  const double *y;
  long nrow;
  std::vector<double> wt;
  bool weighted;
  void raw(const double *x, const double *eta, double *fcost, double *vec) {
    const double c = orc_rtr_raw(P, k, t0, ntiles, y, weighted ? wt.data() : nullptr, x, eta, vec);
    if (fcost) *fcost = c;
  }
  void counts(double *c) { orc_rtr_counts(P, t0, ntiles, c); }
  void unit_weights() { weighted = false; }
  double weights_at(const double *x, double nu, bool keep) {
    if (keep) {
      wt.assign(nrow, 1.0);
      weighted = true;
    }
    return orc_rtr_weights(P, k, t0, ntiles, y, x, nu, keep ? wt.data() : nullptr);
  }
};
}  // namespace

// Y != NULL: consensus terms (rtr_solve_nocuda_robust_admm), kind 5 only
extern "C" void harness_rtr_solve_admm(const void *P, int N, int Nbase, int k, int t0, int ntiles,
                                       const double *y, int kind, double *x, int itmax_a,
                                       int itmax_b, double nulow, double nuhigh, double *robust_nu,
                                       double *info, int nu_joined, const double *Y,
                                       const double *BZ, double rho) {
  CpuEval E;
  E.P = P; E.k = k; E.t0 = t0; E.ntiles = ntiles; E.N = N; E.y = y;
  E.nrow = (long)ntiles * Nbase;
  E.weighted = false;
  rtr::Admm aug = {Y, BZ, rho};
  rtr::solve_chunk(E, kind, x, itmax_a, itmax_b, nulow, nuhigh, robust_nu, info, nu_joined != 0,
                   Y ? &aug : nullptr);
}
extern "C" void harness_rtr_solve(const void *P, int N, int Nbase, int k, int t0, int ntiles,
                                  const double *y, int kind, double *x, int itmax_a, int itmax_b,
                                  double nulow, double nuhigh, double *robust_nu, double *info,
                                  int nu_joined) {
  harness_rtr_solve_admm(P, N, Nbase, k, t0, ntiles, y, kind, x, itmax_a, itmax_b, nulow, nuhigh,
                         robust_nu, info, nu_joined, nullptr, nullptr, 0.0);
}

// nu update of the robust solvers: 1 as meant (default), 0 "sums read before the join are zero"
static int g_nu_joined = 1;
extern "C" void harness_set_nu_joined(int v) { g_nu_joined = v; }

// orc_problem starts with {int N, Nbase, ...}
static void solver_for_oracle(const void *P, int k, int t0, int ntiles, const double *y, int kind,
                              double *x, int itmax_a, int itmax_b, double nulow, double nuhigh,
                              double *robust_nu, double *info) {
  const int *hdr = (const int *)P;
  harness_rtr_solve(P, hdr[0], hdr[1], k, t0, ntiles, y, kind, x, itmax_a, itmax_b, nulow, nuhigh,
                    robust_nu, info, g_nu_joined);
}
__attribute__((constructor)) static void register_solver() { orc_set_rtr_solver(solver_for_oracle); }
