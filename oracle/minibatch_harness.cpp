// TEST INFRASTRUCTURE ONLY (oracle).  The minibatch LBFGS control flow of
// sagecal_b200/csrc/minibatch_algo.h (the code the product runs on the host) on top of the oracle's
// per-row Student's-t cost and gradient (orc_cost / orc_grad, one orc_problem per channel), so that
// it is pinned against the compiled reference's bfgsfit_minibatch_visibilities / _consensus without a
// GPU (tests/test_oracle_minibatch.py).  The product never loads this library.
#include <stdlib.h>
#include <vector>

#include "../sagecal_b200/csrc/minibatch_algo.h"

extern "C" {
double orc_cost(const void *P, const double *pp, const double *x, int robust, double nu);
void orc_grad(const void *P, const double *pp, const double *x, double *g, int robust, double nu);
}

namespace {
struct OracleChan {
  std::vector<const void *> P;
  std::vector<const double *> x;
  int m, Mt, N;
  double nu;
  const double *y, *z, *rho;
  std::vector<double> gtmp;
  double cost(const double *p) {
    double f = 0.0;
    for (size_t c = 0; c < P.size(); c++) f += orc_cost(P[c], p, x[c], 1, nu);
    if (y && z && rho)
      for (int ci = 0; ci < Mt; ci++) {
        double a = 0.0, b = 0.0;
        for (int i = 8 * N * ci; i < 8 * N * (ci + 1); i++) {
          const double xp = p[i] - z[i];
          a += xp * y[i];
          b += xp * xp;
        }
        f += a + rho[ci] * 0.5 * b;
      }
    return f;
  }
  // the reference's multi-channel gradient is MINUS the full-batch one (robust_batchmode_lbfgs.c:1291
  // vs robust_lbfgs.c:299); orc_grad restates the full-batch one
  void grad(const double *p, double *g) {
    for (int i = 0; i < m; i++) g[i] = 0.0;
    for (size_t c = 0; c < P.size(); c++) {
      orc_grad(P[c], p, x[c], gtmp.data(), 1, nu);
      for (int i = 0; i < m; i++) g[i] -= gtmp[i];
    }
    if (y && z && rho)
      for (int ci = 0; ci < Mt; ci++)
        for (int i = 8 * N * ci; i < 8 * N * (ci + 1); i++) g[i] += -y[i] - rho[ci] * (p[i] - z[i]);
  }
};
}  // namespace

extern "C" void *harness_persist_new(int m, int lbfgs_m) {
  persistent_data_t *pt = (persistent_data_t *)calloc(1, sizeof(persistent_data_t));
  pt->s = (double *)calloc((size_t)m * (lbfgs_m + 2) + 8, sizeof(double));
  pt->y = (double *)calloc((size_t)m * lbfgs_m + 1, sizeof(double));
  pt->rho = (double *)calloc((size_t)lbfgs_m + 1, sizeof(double));
  pt->m = m;
  pt->lbfgs_m = lbfgs_m;
  return pt;
}
extern "C" void harness_persist_free(void *p) {
  persistent_data_t *pt = (persistent_data_t *)p;
  free(pt->s); free(pt->y); free(pt->rho); free(pt);
}
extern "C" void harness_minibatch_fit(const void **P, const double **x, int Nf, int N, int Mt,
                                      long nrows, double *p, const double *y, const double *z,
                                      const double *rho, int max_lbfgs, int lbfgs_m, double nu,
                                      double *res_0, double *res_1, void *pt) {
  OracleChan F;
  for (int c = 0; c < Nf; c++) {
    F.P.push_back(P[c]);
    F.x.push_back(x[c]);
  }
  F.m = 8 * N * Mt; F.Mt = Mt; F.N = N; F.nu = nu; F.y = y; F.z = z; F.rho = rho;
  F.gtmp.resize(F.m);
  const double n = (double)nrows * Nf * 8.0;
  *res_0 = F.cost(p);
  minibatch::lbfgs_fit_minibatch(F, p, F.m, max_lbfgs, lbfgs_m, (persistent_data_t *)pt);
  *res_1 = F.cost(p);
  *res_0 /= n;
  *res_1 /= n;
}
