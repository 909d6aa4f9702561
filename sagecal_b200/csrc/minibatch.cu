// Multi-channel minibatch (stochastic) robust LBFGS: bfgsfit_minibatch_visibilities /
// bfgsfit_minibatch_consensus and the persistent state that carries the curvature pairs and the
// on-line gradient variance from one minibatch to the next.
//
// Replaces (reference file:line)
//   bfgsfit_minibatch_visibilities / _consensus     robust_batchmode_lbfgs.c:1446-1577
//   robust_cost_func_multifreq / robust_grad_func_multifreq   robust_batchmode_lbfgs.c:1096-1445
//   lbfgs_fit_minibatch, linesearch_backtrack, mult_hessian   lbfgs.c:717-930, 444-474, 33-111
//   lbfgs_persist_init / _clear / _reset             lbfgs.c:954-1045
//
// Data flow: the reference's data and coherencies are [channel][row][...] arrays with ONE set of Jones
// for all channels of the minibatch.  Every channel becomes a resident single-channel problem (the same
// layout and kernels as the full-batch path: k_stream_all for the Student's-t cost, k_grad_tma_split
// for its gradient), cost and gradient are the sums over the channels, the iterate, the curvature
// pairs and the two-loop recursion (2 M dot products over 8 N Mt doubles) stay on the host like the
// reference's.  Control flow: restated decision for decision; the line search is the reference's
// Armijo backtracking (no numerical differentiation here, unlike the full-batch Fletcher search).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/dirac_b200.h"
#include "problem.h"

// ---- persistent state ---------------------------------------------------------------------------------
// The reference declares persistent_data_t twice (Dirac.h:86-110 CPU build, :196-226 GPU build); the
// two layouts agree only up to `Nt`.  This library touches nothing beyond that common prefix, so a
// caller compiled against either header can hand its struct in: the running averages of the on-line
// gradient variance and the iteration count live behind the curvature pairs in the `s` allocation.
static inline double *pt_running_avg(persistent_data_t *pt) {
  return pt->s + (size_t)pt->m * pt->lbfgs_m;
}
static inline double *pt_running_avg_sq(persistent_data_t *pt) {
  return pt->s + (size_t)pt->m * (pt->lbfgs_m + 1);
}
static inline double *pt_niter(persistent_data_t *pt) {
  return pt->s + (size_t)pt->m * (pt->lbfgs_m + 2);
}

extern "C" int lbfgs_persist_init(persistent_data_t *pt, int Nminibatch, int m, int n, int lbfgs_m,
                                  int Nt) {
  (void)Nminibatch; (void)n;  // the reference's offsets[] / lengths[] tables are never read by its
                              // minibatch drivers (robust_batchmode_lbfgs.c:1493-1495 "not used here")
  const size_t ns = (size_t)m * (lbfgs_m + 2) + 8;
  pt->s = (double *)calloc(ns, sizeof(double));
  pt->y = (double *)calloc((size_t)m * lbfgs_m + 1, sizeof(double));
  pt->rho = (double *)calloc((size_t)lbfgs_m + 1, sizeof(double));
  if (!pt->s || !pt->y || !pt->rho) {
    fprintf(stderr, "%s: %d: no free memory\n", __FILE__, __LINE__);
    exit(1);
  }
  pt->m = m;
  pt->lbfgs_m = lbfgs_m;
  pt->nfilled = 0;
  pt->vacant = 0;
  pt->Nt = Nt;
  return 0;
}
extern "C" int lbfgs_persist_clear(persistent_data_t *pt) {
  free(pt->s);
  free(pt->y);
  free(pt->rho);
  pt->s = pt->y = pt->rho = nullptr;
  return 0;
}
extern "C" int lbfgs_persist_reset(persistent_data_t *pt) {
  memset(pt->s, 0, sizeof(double) * ((size_t)pt->m * (pt->lbfgs_m + 2) + 8));
  memset(pt->y, 0, sizeof(double) * (size_t)pt->m * pt->lbfgs_m);
  memset(pt->rho, 0, sizeof(double) * (size_t)pt->lbfgs_m);
  pt->nfilled = 0;
  pt->vacant = 0;
  return 0;
}

// ---- multi-channel cost / gradient on the device -------------------------------------------------------
struct MultiChan {
  std::vector<dirac_b200_problem *> ch;
  int m, Mt, N;
  double nu;
  const double *y, *z, *rho;  // consensus terms (null: none)
  std::vector<double> gtmp;

  // robust_cost_func_multifreq (robust_batchmode_lbfgs.c:1096-1139)
  double cost(const double *p) {
    double f = 0.0;
    for (auto *pr : ch) f += dirac_b200_predict(pr, p, nullptr, 0, 2, nu);
    if (y && z && rho) {
      for (int ci = 0; ci < Mt; ci++) {
        double a = 0.0, b = 0.0;
        for (int i = 8 * N * ci; i < 8 * N * (ci + 1); i++) {
          const double xp = p[i] - z[i];
          a += xp * y[i];
          b += xp * xp;
        }
        f += a + rho[ci] * 0.5 * b;
      }
    }
    return f;
  }
  // robust_grad_func_multifreq (robust_batchmode_lbfgs.c:1300-1445): sum over the channels of the
  // single-channel Student's-t gradient WITH THE REFERENCE'S SIGN: cpu_calc_deriv_multifreq
  // accumulates -2 sum xr dV / (nu + xr^2) with xr = model - data (:1291), the full-batch
  // cpu_calc_deriv_robust +2 (robust_lbfgs.c:299, the true gradient, which dirac_b200_grad returns).
  // The minibatch LBFGS therefore starts uphill (first step: 2^-15 of the gradient after 15 failed
  // halvings) and only turns once the curvature pairs have negative y^T s; reproduced as is
  // (DESIGN.md 7.8).  The consensus terms enter as the reference writes them (:1420-1438).
  void grad(const double *p, double *g) {
    memset(g, 0, sizeof(double) * m);
    for (auto *pr : ch) {
      dirac_b200_grad(pr, p, gtmp.data(), 1, nu);
      for (int i = 0; i < m; i++) g[i] -= gtmp[i];
    }
    if (y && z && rho)
      for (int ci = 0; ci < Mt; ci++)
        for (int i = 8 * N * ci; i < 8 * N * (ci + 1); i++) g[i] += -y[i] - rho[ci] * (p[i] - z[i]);
  }
};

static double ddot(int m, const double *a, const double *b) {
  double s = 0.0;
  for (int i = 0; i < m; i++) s += a[i] * b[i];
  return s;
}

// pk = H_k gk by the two-loop recursion over the M stored pairs, the newest at slot ii-1
// (mult_hessian, lbfgs.c:33-111)
static void mult_hessian(int m, double *pk, const double *gk, const double *s, const double *y,
                         const double *rho, int M, int ii) {
  std::vector<double> alphai(M > 0 ? M : 1);
  std::vector<int> idx(M > 0 ? M : 1);
  if (M > 0) {
    ii = ii > 0 ? ii - 1 : M - 1;
    for (int ci = 0; ci < M - ii - 1; ci++) idx[ci] = ii + ci + 1;
    for (int ci = M - ii - 1; ci < M; ci++) idx[ci] = ci - M + ii + 1;
  }
  memcpy(pk, gk, sizeof(double) * m);
  for (int ci = 0; ci < M; ci++) {
    const int j = idx[M - ci - 1];
    alphai[M - ci - 1] = rho[j] * ddot(m, s + (size_t)m * j, pk);
    for (int i = 0; i < m; i++) pk[i] -= alphai[M - ci - 1] * y[(size_t)m * j + i];
  }
  if (M > 0) {
    const int j = idx[M - 1];
    const double gamma = ddot(m, s + (size_t)m * j, y + (size_t)m * j) /
                         ddot(m, y + (size_t)m * j, y + (size_t)m * j);
    for (int i = 0; i < m; i++) pk[i] *= gamma;
  }
  for (int ci = 0; ci < M; ci++) {
    const int j = idx[ci];
    const double beta = rho[j] * ddot(m, y + (size_t)m * j, pk);
    for (int i = 0; i < m; i++) pk[i] += (alphai[ci] - beta) * s[(size_t)m * j + i];
  }
}

// Armijo backtracking (linesearch_backtrack, lbfgs.c:444-474)
static double linesearch_backtrack(MultiChan &F, const double *xk, const double *pk,
                                   const double *gk, int m, double alpha0) {
  const double c = 1e-4;
  double alphak = alpha0;
  std::vector<double> xk1(m);
  for (int i = 0; i < m; i++) xk1[i] = xk[i] + alphak * pk[i];
  double fnew = F.cost(xk1.data());
  const double fold = F.cost(xk);
  const double product = c * ddot(m, pk, gk);
  int ci = 0;
  while (ci < 15 && (isnan(fnew) || fnew > fold + alphak * product)) {
    alphak *= 0.5;
    for (int i = 0; i < m; i++) xk1[i] = xk[i] + alphak * pk[i];
    fnew = F.cost(xk1.data());
    ci++;
  }
  return alphak;
}

// lbfgs_fit_minibatch (lbfgs.c:717-930)
static void lbfgs_fit_minibatch(MultiChan &F, double *p, int m, int itmax, int M,
                                persistent_data_t *indata) {
  const double CLM_STOP_THRESH_ = 1e-17, CLM_EPSILON_ = 1e-12;
  std::vector<double> gk(m), xk1(m), xk(p, p + m), pk(m);
  double *s = indata->s, *y = indata->y, *rho = indata->rho;
  double *running_avg = pt_running_avg(indata), *running_avg_sq = pt_running_avg_sq(indata);
  double *niter = pt_niter(indata);
  double alphabar = 1.0;
  F.grad(xk.data(), gk.data());
  double gradnrm = sqrt(ddot(m, gk.data(), gk.data()));
  int ck = gradnrm < CLM_STOP_THRESH_ ? itmax : 0;
  int ci = indata->vacant;
  size_t cm = (size_t)m * ci;
  while (ck < itmax && isnormal(gradnrm) && gradnrm > CLM_STOP_THRESH_) {
    *niter += 1.0;
    const int nit = (int)*niter;
    const bool batch_changed = (nit > 1 && ck == 0);
    if (batch_changed) {
      // running mean / variance of the gradient over the minibatches -> step size cap
      double asum = 0.0;
      for (int i = 0; i < m; i++) {
        const double g_min_rold = gk[i] - running_avg[i];
        running_avg[i] += g_min_rold / (double)nit;
        const double g_min_rnew = gk[i] - running_avg[i];
        running_avg_sq[i] += g_min_rold * g_min_rnew;
      }
      for (int i = 0; i < m; i++) asum += fabs(running_avg_sq[i]);
      alphabar = 10.0 / (1.0 + asum / ((double)(nit - 1) * gradnrm));
    }
    mult_hessian(m, pk.data(), gk.data(), s, y, rho, indata->nfilled < M ? indata->nfilled : M, ci);
    for (int i = 0; i < m; i++) pk[i] = -pk[i];
    const double alphak = linesearch_backtrack(F, xk.data(), pk.data(), gk.data(), m, alphabar);
    if (!isnormal(alphak) || fabs(alphak) < CLM_EPSILON_) break;
    for (int i = 0; i < m; i++) xk1[i] = xk[i] + alphak * pk[i];
    if (!batch_changed)
      for (int i = 0; i < m; i++) {
        s[cm + i] = xk1[i] - xk[i];
        y[cm + i] = -gk[i];
      }
    F.grad(xk1.data(), gk.data());
    gradnrm = sqrt(ddot(m, gk.data(), gk.data()));
    if (!isnormal(gradnrm) || gradnrm < CLM_STOP_THRESH_) break;
    if (!batch_changed) {
      for (int i = 0; i < m; i++) y[cm + i] += gk[i];
      const double lm0 = 1e-6;
      if (gradnrm > 1e3 * lm0)
        for (int i = 0; i < m; i++) y[cm + i] += lm0 * s[cm + i];
      rho[ci] = 1.0 / ddot(m, y + cm, s + cm);
    }
    xk = xk1;
    ck++;
    if (!batch_changed) {
      indata->nfilled = (indata->nfilled < M ? indata->nfilled + 1 : M);
      if (cm < (size_t)(M - 1) * m) {
        cm += m;
        ci++;
        indata->vacant++;
      } else {
        cm = 0;
        ci = 0;
        indata->vacant = 0;
      }
    }
  }
  memcpy(p, xk.data(), sizeof(double) * m);
}

static int minibatch_fit(double *x, int N, int Nbase, int tilesz, baseline_t *barr,
                         clus_source_t *carr, double *coh, int M, int Mt, int Nf, double *p,
                         const double *y, const double *z, const double *rho, int max_lbfgs,
                         int lbfgs_m, double robust_nu, double *res_0, double *res_1,
                         persistent_data_t *indata) {
  const int m = N * Mt * 8;
  const long long R = (long long)Nbase * tilesz;
  const double n = (double)R * Nf * 8.0;
  MultiChan F;
  F.m = m; F.Mt = Mt; F.N = N; F.nu = robust_nu; F.y = y; F.z = z; F.rho = rho;
  F.gtmp.resize(m);
  // channel c: coh[c][row][M][4] (complex), x[c][row][8]  (robust_batchmode_lbfgs.c:1176-1183)
  for (int c = 0; c < Nf; c++)
    F.ch.push_back(dirac_b200_create(N, Nbase, tilesz, barr, carr, M, Mt,
                                     coh + (size_t)c * 8 * M * R, x + (size_t)c * 8 * R));
  *res_0 = F.cost(p);
  // lbfgs_fit (lbfgs.c:933-950): persistent data -> minibatch variant
  lbfgs_fit_minibatch(F, p, m, max_lbfgs, lbfgs_m, indata);
  *res_1 = F.cost(p);
  *res_0 *= 1.0 / n;
  *res_1 *= 1.0 / n;
  for (auto *pr : F.ch) dirac_b200_destroy(pr);
  return 0;
}

extern "C" int bfgsfit_minibatch_visibilities(double *u, double *v, double *w, double *x, int N,
                                              int Nbase, int tilesz, baseline_t *barr,
                                              clus_source_t *carr, double *coh, int M, int Mt,
                                              double *freqs, int Nf, double fdelta, double *p, int Nt,
                                              int max_lbfgs, int lbfgs_m, int gpu_threads,
                                              int solver_mode, double robust_nu, double *res_0,
                                              double *res_1, persistent_data_t *indata,
                                              int nminibatch, int totalminibatch) {
  (void)u; (void)v; (void)w; (void)freqs; (void)fdelta; (void)Nt; (void)gpu_threads;
  (void)solver_mode; (void)nminibatch; (void)totalminibatch;
  return minibatch_fit(x, N, Nbase, tilesz, barr, carr, coh, M, Mt, Nf, p, nullptr, nullptr, nullptr,
                       max_lbfgs, lbfgs_m, robust_nu, res_0, res_1, indata);
}

extern "C" int bfgsfit_minibatch_consensus(double *u, double *v, double *w, double *x, int N,
                                           int Nbase, int tilesz, baseline_t *barr,
                                           clus_source_t *carr, double *coh, int M, int Mt,
                                           double *freqs, int Nf, double fdelta, double *p, double *y,
                                           double *z, double *rho, int Nt, int max_lbfgs, int lbfgs_m,
                                           int gpu_threads, int solver_mode, double robust_nu,
                                           double *res_0, double *res_1, persistent_data_t *indata,
                                           int nminibatch, int totalminibatch) {
  (void)u; (void)v; (void)w; (void)freqs; (void)fdelta; (void)Nt; (void)gpu_threads;
  (void)solver_mode; (void)nminibatch; (void)totalminibatch;
  return minibatch_fit(x, N, Nbase, tilesz, barr, carr, coh, M, Mt, Nf, p, y, z, rho, max_lbfgs,
                       lbfgs_m, robust_nu, res_0, res_1, indata);
}
