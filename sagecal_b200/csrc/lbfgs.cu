// LBFGS (two-loop recursion + Fletcher line search) on top of the device cost / gradient passes.
//
// The iteration logic restates lbfgs_fit_fullbatch (lbfgs.c:479-640), mult_hessian (:33-111),
// linesearch (:298-430), linesearch_zoom (:211-290) and cubic_interp (:116-205): same constants,
// same order of cost evaluations, same acceptance tests, because every comparison steers the
// iterates and parity with the CPU reference is judged on the solved Jones.  The parameter vector
// is small (8*N*Mt doubles); it lives on the host like in the reference, while every cost/gradient
// evaluation is one or two streaming passes over the resident coherencies:
//   cost  = k_predict_full (sum e^2 | sum log(1+e^2/nu))        replaces cost_func / robust_cost_func
//   grad  = k_predict_full (residual) + k_grad_full             replaces func_grad(_robust)
#include <float.h>
#include <math.h>
#include <string.h>
#include <vector>

#include "problem.h"

struct LbfgsCtx {
  dirac_b200_problem *pr;
  int robust;
  double nu;
  long long ncost, ngrad;
};

static double cost_eval(LbfgsCtx *c, const double *p, int m) {
  DevProblem &d = c->pr->d;
  DB_CHECK(cudaMemcpyAsync(d.pp, p, sizeof(double) * m, cudaMemcpyHostToDevice, d.stream));
  db_predict_dev(c->pr, d.pp, nullptr, 0, c->robust ? 2 : 1, c->nu, 0);
  c->ncost++;
  return db_read_scalar(c->pr, 0);
}

static void grad_eval(LbfgsCtx *c, const double *p, double *g, int m) {
  DevProblem &d = c->pr->d;
  DB_CHECK(cudaMemcpyAsync(d.pp, p, sizeof(double) * m, cudaMemcpyHostToDevice, d.stream));
  db_predict_dev(c->pr, d.pp, c->pr->res, 1, 0, 0.0, 0);
  db_grad_dev(c->pr, d.pp, c->pr->g, c->robust, c->nu);
  DB_CHECK(cudaMemcpyAsync(g, c->pr->g, sizeof(double) * m, cudaMemcpyDeviceToHost, d.stream));
  DB_CHECK(cudaStreamSynchronize(d.stream));
  c->ngrad++;
}

static inline double vdot(const double *a, const double *b, int m) {
  double s = 0.0;
  for (int i = 0; i < m; i++) s += a[i] * b[i];
  return s;
}
static inline void vaxpy(double *y, const double *x, double a, int m) {  // y += a x
  for (int i = 0; i < m; i++) y[i] += a * x[i];
}
static inline double vnrm2(const double *a, int m) { return sqrt(vdot(a, a, m)); }

// pk = H_k gk by the two-loop recursion; M = number of valid pairs, ii = slot to be written next
static void mult_hessian(int m, double *pk, const double *gk, const double *s, const double *y,
                         const double *rho, int M, int ii) {
  std::vector<double> alphai(M > 0 ? M : 1);
  std::vector<int> idx(M > 0 ? M : 1);
  if (M > 0) {
    ii = (ii > 0) ? ii - 1 : M - 1;
    for (int ci = 0; ci < M - ii - 1; ci++) idx[ci] = ii + ci + 1;
    for (int ci = M - ii - 1; ci < M; ci++) idx[ci] = ci - M + ii + 1;
  }
  memcpy(pk, gk, sizeof(double) * m);
  for (int ci = 0; ci < M; ci++) {
    int j = idx[M - ci - 1];
    alphai[M - ci - 1] = rho[j] * vdot(&s[(size_t)m * j], pk, m);
    vaxpy(pk, &y[(size_t)m * j], -alphai[M - ci - 1], m);
  }
  if (M > 0) {
    int j = idx[M - 1];
    double gamma = vdot(&s[(size_t)m * j], &y[(size_t)m * j], m);
    gamma /= vdot(&y[(size_t)m * j], &y[(size_t)m * j], m);
    for (int i = 0; i < m; i++) pk[i] *= gamma;
  }
  for (int ci = 0; ci < M; ci++) {
    int j = idx[ci];
    double beta = rho[j] * vdot(&y[(size_t)m * j], pk, m);
    vaxpy(pk, &s[(size_t)m * j], alphai[ci] - beta, m);
  }
}

static double cubic_interp(LbfgsCtx *c, const double *xk, const double *pk, double a, double b,
                           double *xp, int m, double step) {
  double f0, f1, f0d, f1d, p01, p02, z0, fz0, aa, cc;
  memcpy(xp, xk, sizeof(double) * m);
  vaxpy(xp, pk, a, m);
  f0 = cost_eval(c, xp, m);
  vaxpy(xp, pk, step, m);
  p01 = cost_eval(c, xp, m);
  vaxpy(xp, pk, -2.0 * step, m);
  p02 = cost_eval(c, xp, m);
  f0d = (p01 - p02) / (2.0 * step);
  vaxpy(xp, pk, -a + step + b, m);
  f1 = cost_eval(c, xp, m);
  vaxpy(xp, pk, step, m);
  p01 = cost_eval(c, xp, m);
  vaxpy(xp, pk, -2.0 * step, m);
  p02 = cost_eval(c, xp, m);
  f1d = (p01 - p02) / (2.0 * step);

  aa = 3.0 * (f0 - f1) / (b - a) + (f1d - f0d);
  p01 = aa * aa - f0d * f1d;
  if (p01 > 0.0) {
    cc = sqrt(p01);
    z0 = b - (f1d + cc - aa) * (b - a) / (f1d - f0d + 2.0 * cc);
    aa = (a > b) ? a : b;
    cc = (a < b) ? a : b;
    if (z0 > aa || z0 < cc) {
      fz0 = f0 + f1;
    } else {
      vaxpy(xp, pk, -b + step + a + z0 * (b - a), m);
      fz0 = cost_eval(c, xp, m);
    }
    if (f0 < f1 && f0 < fz0) return a;
    if (f1 < fz0) return b;
    return z0;
  }
  return (f0 < f1) ? a : b;
}

static double linesearch_zoom(LbfgsCtx *c, const double *xk, const double *pk, double a, double b,
                              double *xp, double phi_0, double gphi_0, double sigma, double rho,
                              double t1, double t2, double t3, int m, double step) {
  double alphaj = 0.0, phi_j, phi_aj, gphi_j, p01, p02, aj = a, bj = b, alphak = 1.0;
  int ci = 0, found_step = 0;
  (void)t1;
  while (ci < 10) {
    p01 = aj + t2 * (bj - aj);
    p02 = bj - t3 * (bj - aj);
    alphaj = cubic_interp(c, xk, pk, p01, p02, xp, m, step);
    memcpy(xp, xk, sizeof(double) * m);
    vaxpy(xp, pk, alphaj, m);
    phi_j = cost_eval(c, xp, m);
    vaxpy(xp, pk, -alphaj + aj, m);
    phi_aj = cost_eval(c, xp, m);
    if ((phi_j > phi_0 + rho * alphaj * gphi_0) || phi_j >= phi_aj) {
      bj = alphaj;
    } else {
      vaxpy(xp, pk, -aj + alphaj + step, m);
      p01 = cost_eval(c, xp, m);
      vaxpy(xp, pk, -2.0 * step, m);
      p02 = cost_eval(c, xp, m);
      gphi_j = (p01 - p02) / (2.0 * step);
      if ((aj - alphaj) * gphi_j <= step) {
        alphak = alphaj;
        found_step = 1;
        break;
      }
      if (fabs(gphi_j) <= -sigma * gphi_0) {
        alphak = alphaj;
        found_step = 1;
        break;
      }
      if (gphi_j * (bj - aj) >= 0) bj = aj;
      aj = alphaj;
    }
    ci++;
  }
  if (!found_step) alphak = alphaj;
  return alphak;
}

static double linesearch(LbfgsCtx *c, const double *xk, const double *pk, double alpha1,
                         double sigma, double rho, double t1, double t2, double t3, int m,
                         double step) {
  std::vector<double> xpv(m);
  double *xp = xpv.data();
  double alphai, alphai1, phi_0, phi_alphai, phi_alphai1, p01, p02, gphi_0, gphi_i, alphak, mu, tol;
  alphak = 1.0;
  phi_0 = cost_eval(c, xk, m);
  tol = (0.01 * phi_0 < 1e-6) ? 0.01 * phi_0 : 1e-6;
  memcpy(xp, xk, sizeof(double) * m);
  vaxpy(xp, pk, step, m);
  p01 = cost_eval(c, xp, m);
  vaxpy(xp, pk, -2.0 * step, m);
  p02 = cost_eval(c, xp, m);
  gphi_0 = (p01 - p02) / (2.0 * step);
  mu = (tol - phi_0) / (rho * gphi_0);
  if (!isnormal(mu)) return mu;
  int ci = 1;
  alphai = alpha1;
  alphai1 = 0.0;
  phi_alphai1 = phi_0;
  while (ci < 10) {
    memcpy(xp, xk, sizeof(double) * m);
    vaxpy(xp, pk, alphai, m);
    phi_alphai = cost_eval(c, xp, m);
    if (phi_alphai < tol) {
      alphak = alphai;
      break;
    }
    if ((phi_alphai > phi_0 + alphai * gphi_0) || (ci > 1 && phi_alphai >= phi_alphai1)) {
      alphak = linesearch_zoom(c, xk, pk, alphai1, alphai, xp, phi_0, gphi_0, sigma, rho, t1, t2,
                               t3, m, step);
      break;
    }
    vaxpy(xp, pk, step, m);
    p01 = cost_eval(c, xp, m);
    vaxpy(xp, pk, -2.0 * step, m);
    p02 = cost_eval(c, xp, m);
    gphi_i = (p01 - p02) / (2.0 * step);
    if (fabs(gphi_i) <= -sigma * gphi_0) {
      alphak = alphai;
      break;
    }
    if (gphi_i >= 0) {
      alphak = linesearch_zoom(c, xk, pk, alphai, alphai1, xp, phi_0, gphi_0, sigma, rho, t1, t2,
                               t3, m, step);
      break;
    }
    if (mu <= (2.0 * alphai - alphai1)) {
      alphai1 = alphai;
      alphai = mu;
    } else {
      p01 = 2.0 * alphai - alphai1;
      double hi = alphai + t1 * (alphai - alphai1);
      p02 = (mu < hi) ? mu : hi;
      alphai = cubic_interp(c, xk, pk, p01, p02, xp, m, step);
    }
    phi_alphai1 = phi_alphai;
    ci++;
  }
  return alphak;
}

// p: m x 1 in/out (host).  robust != 0 -> Student's-t cost with nu.
void db_lbfgs_fit(dirac_b200_problem *pr, double *p, int m, int itmax, int M, int robust,
                  double nu) {
  LbfgsCtx ctx;
  ctx.pr = pr;
  ctx.robust = robust;
  ctx.nu = nu;
  ctx.ncost = ctx.ngrad = 0;
  if (M < 1) M = 1;
  std::vector<double> gk(m), xk1(m), xk(m), pk(m), s((size_t)m * M), y((size_t)m * M), rho(M);
  double step, alphak;
  int ck, ci, cm;
  memcpy(xk.data(), p, sizeof(double) * m);
  grad_eval(&ctx, xk.data(), gk.data(), m);
  double gradnrm = vnrm2(gk.data(), m);
  const double STOP = 1e-17;  // CLM_STOP_THRESH, Dirac_common.h:43
  if (gradnrm < STOP) {
    ck = itmax;
    step = 0.0;
  } else {
    ck = 0;
    double t = 1e-3 / gradnrm;
    if (t > 1e-6) t = 1e-6;
    step = (t > 1e-9) ? t : 1e-9;
  }
  cm = 0;
  ci = 0;
  while (ck < itmax && isnormal(gradnrm) && gradnrm > STOP) {
    mult_hessian(m, pk.data(), gk.data(), s.data(), y.data(), rho.data(), ck < M ? ck : M, ci);
    for (int i = 0; i < m; i++) pk[i] = -pk[i];
    alphak = linesearch(&ctx, xk.data(), pk.data(), 10.0, 0.1, 0.01, 9, 0.1, 0.5, m, step);
    if (!isnormal(alphak) || fabs(alphak) < 1e-12) break;  // CLM_EPSILON
    memcpy(xk1.data(), xk.data(), sizeof(double) * m);
    vaxpy(xk1.data(), pk.data(), alphak, m);
    double *sk = &s[(size_t)cm];
    double *yk = &y[(size_t)cm];
    for (int i = 0; i < m; i++) {
      sk[i] = xk1[i] - xk[i];
      yk[i] = -gk[i];
    }
    grad_eval(&ctx, xk1.data(), gk.data(), m);
    gradnrm = vnrm2(gk.data(), m);
    vaxpy(yk, gk.data(), 1.0, m);
    rho[ci] = 1.0 / vdot(yk, sk, m);
    memcpy(xk.data(), xk1.data(), sizeof(double) * m);
    ck++;
    if (cm < (M - 1) * m) {
      cm += m;
      ci++;
    } else {
      cm = ci = 0;
    }
  }
  memcpy(p, xk.data(), sizeof(double) * m);
}
