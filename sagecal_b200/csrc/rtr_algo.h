// Riemannian trust-region (RTR), Riemannian steepest-descent (RSD) and Nesterov (NSD) solvers of one
// (cluster, chunk): the control flow, as host code templated on the evaluator that supplies the raw
// per-station sums.  The product instantiates it with the device evaluator (rtr.cu: k_rtr_stats /
// k_rtr_eval); tests/rtr_harness.cpp instantiates it with a plain O(rows) CPU evaluator to pin the
// control flow against the compiled reference without a GPU (test infrastructure, not shipped).
//
// Restatement, decision for decision, of
//   rtr_solve_nocuda          rtr_solve.c:1207-1610   (RSD warm-up by Armijo steps + RTR)
//   rtr_solve_nocuda_robust   rtr_solve_robust.c:1440-1875 (initial radius by Sartenaer's ITRR,
//                             Student's-t row weights, RTR on the weighted cost)
//   nsd_solve_nocuda_robust   rtr_solve_robust.c:1877-2136 (Nesterov's accelerated descent)
//   tcg_solve / armijostep / itrr / fns_proj / fns_g (rtr_solve.c:886-1204, rtr_solve_robust.c:1308-1437)
// including their quirks (flagged in DESIGN.md): the Hessian is NOT scaled by the per-station
// baseline counts although the gradient is (the scaling loop runs over the output buffer before
// the projection overwrites it, rtr_solve.c:847-866), tcg_solve adds the caller's `fhess` buffer
// to H.delta (rtr_solve.c:991,1016: zero in the plain solver, the ITRR's stale Hessian in the robust
// one), the robust solver compares the unweighted entry cost with weighted trial costs
// (rtr_solve_robust.c:1575,1607,1700), NSD never sets info[0].
//
// Vectors are 8N doubles in the API's parameter layout: station s holds J[a][m] at 8s + 2(2a+m)
// (re, im); the reference's 2N x 2 complex layout (rtr_solve.c:1224-1243) is a permutation of it and
// every operation below is layout independent except the projection, which is written out.
//
// Evaluator concept:
//   int N;                                         stations
//   void raw(const double *x, const double *eta, double *fcost, double *vec);
//        fcost != null: sum_rows w |d - Jp C Jq^H|^2 ; vec != null: per-station sums of
//        w res Jq C^H / w res^H Jp C (eta == null) or of the Hessian terms (eta != null)
//   void counts(double *c);                        unflagged rows per station
//   void unit_weights();                           w = 1 from here on
//   double weights_at(const double *x, double nu, bool keep);
//        row weights (nu+2)/(nu+max_c|res_c|^2) at x; returns mean over ALL rows of log w - w;
//        keep: the weights apply from here on
#pragma once
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace rtr {

// complex helpers on the parameter layout: station s holds J[a][m] at 8s + 2(2a+m) (re, im)
inline double rtr_g(int n8, const double *a, const double *b) {  // fns_g: 2 Re tr(a^H b)
  double s = 0.0;
  for (int i = 0; i < n8; i++) s += a[i] * b[i];
  return 2.0 * s;
}
inline double rtr_nrm2(int n8, const double *a) {
  double s = 0.0;
  for (int i = 0; i < n8; i++) s += a[i] * a[i];
  return sqrt(s);
}
inline void rtr_axpy(int n8, const double *x, double a, double *y) {
  for (int i = 0; i < n8; i++) y[i] += a * x[i];
}

struct cplx {
  double r, i;
};
inline cplx cm(cplx a, cplx b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
inline cplx cj(cplx a) { return {a.r, -a.i}; }
inline cplx cs(cplx a, cplx b) { return {a.r - b.r, a.i - b.i}; }
inline cplx ca(cplx a, cplx b) { return {a.r + b.r, a.i + b.i}; }
inline cplx cdivc(cplx a, cplx b) {
  const double dd = b.r * b.r + b.i * b.i;
  return {(a.r * b.r + a.i * b.i) / dd, (a.i * b.r - a.r * b.i) / dd};
}

// fns_proj (rtr_solve.c:339-410): Z - X Om with Om X^H X + X^H X Om = X^H Z - Z^H X
inline void proj(int N, const double *x, const double *z, double *out) {
  // columns of X (2N x 2): column m holds J_s[a][m] for all s, a
  cplx xx[2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}}, xz[2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
  for (int s = 0; s < N; s++)
    for (int a = 0; a < 2; a++)
      for (int m = 0; m < 2; m++) {
        const cplx xm = {x[8 * s + 2 * (2 * a + m)], x[8 * s + 2 * (2 * a + m) + 1]};
        for (int mp = 0; mp < 2; mp++) {
          const cplx xp = {x[8 * s + 2 * (2 * a + mp)], x[8 * s + 2 * (2 * a + mp) + 1]};
          const cplx zp = {z[8 * s + 2 * (2 * a + mp)], z[8 * s + 2 * (2 * a + mp) + 1]};
          xx[m][mp] = ca(xx[m][mp], cm(cj(xm), xp));
          xz[m][mp] = ca(xz[m][mp], cm(cj(xm), zp));
        }
      }
  const cplx xx00 = xx[0][0], xx01 = xx[0][1], xx10 = cj(xx[0][1]), xx11 = xx[1][1];
  const cplx rr00 = cs(xz[0][0], cj(xz[0][0]));
  const cplx rr01 = cs(xz[0][1], cj(xz[1][0]));
  const cplx rr10 = {-rr01.r, rr01.i};  // -conj(rr01)
  const cplx rr11 = cs(xz[1][1], cj(xz[1][1]));
  // A u = b, u = vec(Om) column-major (rtr_solve.c:365-381); Gaussian elimination, partial pivoting
  cplx A[4][5];
  const cplx zero = {0, 0};
  const cplx d01 = ca(xx11, xx00);
  A[0][0] = {2.0 * xx00.r, 2.0 * xx00.i}; A[0][1] = xx01; A[0][2] = xx10; A[0][3] = zero;
  A[1][0] = xx10; A[1][1] = d01; A[1][2] = zero; A[1][3] = xx10;
  A[2][0] = xx01; A[2][1] = zero; A[2][2] = d01; A[2][3] = xx01;
  A[3][0] = zero; A[3][1] = xx01; A[3][2] = xx10; A[3][3] = {2.0 * xx11.r, 2.0 * xx11.i};
  A[0][4] = rr00; A[1][4] = rr10; A[2][4] = rr01; A[3][4] = rr11;
  cplx u[4] = {zero, zero, zero, zero};
  bool singular = false;
  for (int c = 0; c < 4; c++) {
    int piv = c;
    double best = A[c][c].r * A[c][c].r + A[c][c].i * A[c][c].i;
    for (int r = c + 1; r < 4; r++) {
      const double v = A[r][c].r * A[r][c].r + A[r][c].i * A[r][c].i;
      if (v > best) { best = v; piv = r; }
    }
    if (best == 0.0) { singular = true; break; }
    if (piv != c)
      for (int j = 0; j < 5; j++) std::swap(A[c][j], A[piv][j]);
    for (int r = c + 1; r < 4; r++) {
      const cplx fct = cdivc(A[r][c], A[c][c]);
      for (int j = c; j < 5; j++) A[r][j] = cs(A[r][j], cm(fct, A[c][j]));
    }
  }
  if (!singular) {
    for (int r = 3; r >= 0; r--) {
      cplx acc = A[r][4];
      for (int j = r + 1; j < 4; j++) acc = cs(acc, cm(A[r][j], u[j]));
      u[r] = cdivc(acc, A[r][r]);
    }
  }
  // Om[m'][m] = u[m' + 2m];  out_s[a][m] = z_s[a][m] - sum_m' x_s[a][m'] Om[m'][m]
  for (int s = 0; s < N; s++)
    for (int a = 0; a < 2; a++)
      for (int m = 0; m < 2; m++) {
        cplx v = {z[8 * s + 2 * (2 * a + m)], z[8 * s + 2 * (2 * a + m) + 1]};
        for (int mp = 0; mp < 2; mp++) {
          const cplx xp = {x[8 * s + 2 * (2 * a + mp)], x[8 * s + 2 * (2 * a + mp) + 1]};
          v = cs(v, cm(xp, u[mp + 2 * m]));
        }
        out[8 * s + 2 * (2 * a + m)] = v.r;
        out[8 * s + 2 * (2 * a + m) + 1] = v.i;
      }
}


// digamma (updatenu.c:36-49)
inline double digamma(double x) {
  double result = 0.0, xx, xx2, xx4;
  for (; x < 7.0; ++x) result -= 1.0 / x;
  x -= 0.5;
  xx = 1.0 / x;
  xx2 = xx * xx;
  xx4 = xx2 * xx2;
  result += log(x) + (1. / 24.) * xx2 - (7.0 / 960.0) * xx4 + (31.0 / 8064.0) * xx4 * xx2 -
            (127.0 / 30720.0) * xx4 * xx4;
  return result;
}

// consensus (ADMM) terms of one (cluster, chunk): cost + y^H (J - BZ) + rho/2 |J - BZ|^2
// (rtr_solve_robust_admm.c:199-214); Y, BZ: 8N doubles in the parameter layout
struct Admm {
  const double *Y, *BZ;
  double rho;
};

// the fns_* callbacks of the reference on top of an evaluator's raw sums.  With consensus terms
// (rtr_solve_robust_admm.c) the search space is Euclidean: the projection is the identity (:425-430)
template <class EV>
struct Ops {
  EV &ev;
  int N, n8;
  const Admm *aug;
  std::vector<double> iw;   // per-station inverse baseline counts, max 1 (fns_fcount)
  std::vector<double> rawv;
  explicit Ops(EV &e, const Admm *a = nullptr)
      : ev(e), N(e.N), n8(8 * e.N), aug(a), iw(e.N, 0.0), rawv(8 * e.N, 0.0) {}
  void project(const double *x, const double *z, double *out) const {
    if (aug) memcpy(out, z, sizeof(double) * n8);
    else proj(N, x, z, out);
  }
  // 2 Re(Y^H (x - BZ)) + rho/2 |x - BZ|^2  (rtr_solve_robust_admm.c:205-212)
  double aug_cost(const double *x) const {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < n8; i++) {
      const double dlt = x[i] - aug->BZ[i];
      a += aug->Y[i] * dlt;
      b += dlt * dlt;
    }
    return 2.0 * a + 0.5 * aug->rho * b;
  }
  // fns_fcount (rtr_solve.c:99-180): inverse of the unflagged rows per station, scaled to max 1
  void count() {
    std::vector<double> c(N);
    ev.counts(c.data());
    double mx = 0.0;
    for (int i = 0; i < N; i++) {
      iw[i] = c[i] > 0.0 ? 1.0 / c[i] : 0.0;
      if (fabs(iw[i]) > mx) mx = fabs(iw[i]);
    }
    if (mx > 0.0)
      for (int i = 0; i < N; i++) iw[i] *= 1.0 / mx;
  }
  double f(const double *x) {  // fns_f (rtr_solve.c:251, rtr_solve_robust.c:135)
    double c;
    ev.raw(x, nullptr, &c, nullptr);
    if (aug) c += aug_cost(x);
    return c;
  }
  // fns_fgrad (rtr_solve.c:539-636): station sums scaled by iw, optionally negated, projected.
  // fx != null: the cost at x rides along in the same evaluation.
  void fgrad(const double *x, double *g, bool negate, double *fx = nullptr) {
    ev.raw(x, nullptr, fx, rawv.data());
    if (fx && aug) *fx += aug_cost(x);
    for (int s = 0; s < N; s++) {
      const double sc = negate ? -iw[s] : iw[s];
      for (int i = 0; i < 8; i++) rawv[8 * s + i] *= sc;
    }
    if (aug) {  // +-(Y/2 + rho/2 (x - BZ))  (rtr_solve_robust_admm.c:679-687)
      const double sg = negate ? 0.5 : -0.5;
      for (int i = 0; i < n8; i++)
        rawv[i] += sg * (aug->Y[i] + aug->rho * (x[i] - aug->BZ[i]));
    }
    project(x, rawv.data(), g);
  }
  // fns_fhess (rtr_solve.c:774-870): NOT scaled by iw (see the file header), projected
  void fhess(const double *x, const double *eta, double *h) {
    ev.raw(x, eta, nullptr, rawv.data());
    if (aug)  // + rho/2 eta  (rtr_solve_robust_admm.c:949)
      for (int i = 0; i < n8; i++) rawv[i] += 0.5 * aug->rho * eta[i];
    project(x, rawv.data(), h);
  }
};

// ------------------------------------------------------------------------------------------------
// truncated CG on the tangent space (tcg_solve, rtr_solve.c:886-1145).  eta: in (zero) / out.
// fhess: the CALLER's buffer that the reference adds to H.delta (see the file header).
// ------------------------------------------------------------------------------------------------
template <class EV>
static int tcg_solve(Ops<EV> &E, const double *x, const double *grad, double *eta,
                     const double *fhess, double Delta, double theta, double kappa, int max_inner,
                     int min_inner) {
  const int n = E.n8;
  std::vector<double> r(grad, grad + n), z(n), delta(n), Hxd(n);
  double e_Pe = 0.0;
  double r_r = rtr_g(n, r.data(), r.data());
  double norm_r = sqrt(r_r);
  const double norm_r0 = norm_r;
  z = r;
  double z_r = rtr_g(n, z.data(), r.data());
  double d_Pd = z_r;
  for (int i = 0; i < n; i++) delta[i] = -z[i];
  double e_Pd = rtr_g(n, eta, delta.data());
  int stop_tCG = 5;
  const double Deltasq = Delta * Delta;
  for (int cj_ = 1; cj_ <= max_inner; cj_++) {
    E.fhess(x, delta.data(), Hxd.data());
    const double d_Hd = rtr_g(n, delta.data(), Hxd.data());
    const double alpha = z_r / d_Hd;
    const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
    if (d_Hd <= 0.0 || e_Pe_new >= Deltasq) {
      const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Deltasq - e_Pe))) / d_Pd;
      rtr_axpy(n, delta.data(), tau, eta);
      rtr_axpy(n, fhess, tau, Hxd.data());
      stop_tCG = (d_Hd <= 0.0 ? 1 : 2);
      break;
    }
    e_Pe = e_Pe_new;
    rtr_axpy(n, delta.data(), alpha, eta);
    rtr_axpy(n, fhess, alpha, Hxd.data());
    rtr_axpy(n, Hxd.data(), alpha, r.data());
    r_r = rtr_g(n, r.data(), r.data());
    norm_r = sqrt(r_r);
    if (cj_ >= min_inner) {
      const double norm_r0pow = pow(norm_r0, theta);
      if (norm_r <= norm_r0 * std::min(norm_r0pow, kappa)) {
        stop_tCG = (kappa < norm_r0pow ? 3 : 4);
        break;
      }
    }
    z = r;
    const double zold_rold = z_r;
    z_r = rtr_g(n, z.data(), r.data());
    const double beta = z_r / zold_rold;
    for (int i = 0; i < n; i++) delta[i] = beta * delta[i] - z[i];
    e_Pd = beta * (e_Pd + alpha * d_Pd);
    d_Pd = z_r + beta * beta * d_Pd;
  }
  return stop_tCG;
}

// Armijo step of the RSD warm-up (armijostep, rtr_solve.c:1156-1204): teta out; returns 1 if the
// cost did not go down
template <class EV>
static int armijostep(Ops<EV> &E, const double *x, double *teta, double *eta, double *x_prop,
                      double *mincost) {
  const int n = E.n8;
  const double alphabar = 10.0, beta = 0.2, sigma = 0.5;
  double fx;
  E.fgrad(x, eta, false, &fx);
  double beta0 = beta;
  double minfx = fx, minbeta = beta0;
  double lhs = fx, rhs;
  int nocostred = 0;
  *mincost = fx;
  const double metric0 = rtr_g(n, eta, eta);
  for (int m = 0; m < 50; m++) {
    for (int i = 0; i < n; i++) {
      teta[i] = eta[i] * (beta0 * alphabar);
      x_prop[i] = x[i] + teta[i];
    }
    lhs = E.f(x_prop);
    if (lhs < minfx) {
      minfx = lhs;
      *mincost = minfx;
      minbeta = beta0;
    }
    const double metric = beta0 * alphabar * metric0;
    rhs = fx + sigma * metric;
    if (lhs <= rhs) {
      minbeta = beta0;
      break;
    }
    beta0 = beta0 * beta;
  }
  if (lhs > fx) nocostred = 1;
  for (int i = 0; i < n; i++) teta[i] = eta[i] * (minbeta * alphabar);
  return nocostred;
}

// trust-region loop shared by the plain and the robust solver (rtr_solve.c:1370-1570,
// rtr_solve_robust.c:1644-1837).  fx in/out, x in/out.
template <class EV>
static void tr_loop(Ops<EV> &E, double *x, double &fx, double *fgradx, double *eta, double *Heta,
                    double *x_prop, int itmax_rtr, double Delta_bar, double Delta0,
                    double rho_regularization) {
  const int n = E.n8;
  const int min_inner = 1, max_inner = itmax_rtr, min_outer = 3, max_outer = itmax_rtr;
  const double epsilon = 1e-12 /* CLM_EPSILON */, kappa = 0.1, theta = 1.0;
  const double eta1 = 0.0001, eta2 = 0.99, alpha1 = 0.25, alpha2 = 3.5;
  const double rho_prime = eta1;
  int k = 0;
  int stop_outer = (itmax_rtr > 0 ? 0 : 1);
  double norm_grad = 0.0;
  std::vector<double> gprop(n);
  if (!stop_outer) {
    E.fgrad(x, fgradx, true);
    norm_grad = sqrt(rtr_g(n, fgradx, fgradx));
  }
  double Delta = Delta0;
  while (!stop_outer) {
    k++;
    memset(eta, 0, sizeof(double) * n);
    const int stop_inner =
        tcg_solve(E, x, fgradx, eta, Heta, Delta, theta, kappa, max_inner, min_inner);
    for (int i = 0; i < n; i++) x_prop[i] = x[i] + eta[i];
    // the gradient at the proposal rides along with its cost (one evaluation instead of two when the
    // step is accepted, which is the usual case; the reference evaluates them separately,
    // rtr_solve.c:1436,1530)
    double fx_prop;
    E.fgrad(x_prop, gprop.data(), true, &fx_prop);
    double rhonum = fx - fx_prop;
    double rhoden = -rtr_g(n, fgradx, eta) - 0.5 * rtr_g(n, Heta, eta);
    const double rho_reg = std::max(1.0, fx) * rho_regularization;
    rhonum += rho_reg;
    rhoden += rho_reg;
    const double rho = rhonum / rhoden;
    const int model_decreased = (rhoden >= 0.0 ? 1 : 0);
    if (!model_decreased || rho < eta1) {
      Delta = alpha1 * Delta;
    } else if (rho > eta2 && (stop_inner == 2 || stop_inner == 1)) {
      Delta = std::min(alpha2 * Delta, Delta_bar);
    }
    if (model_decreased && rho > rho_prime) {
      memcpy(x, x_prop, sizeof(double) * n);
      fx = fx_prop;
      memcpy(fgradx, gprop.data(), sizeof(double) * n);
      norm_grad = sqrt(rtr_g(n, fgradx, fgradx));
    }
    if (norm_grad < epsilon && k > min_outer) stop_outer = 1;
    if (k >= max_outer) stop_outer = 1;
  }
}

// update_nu with p = 2 (AECM, updatenu.c:262-340)
inline double update_nu_aecm(double logsumw, double nulow, double nuhigh, int p, double nu_old) {
  const int Nd = 30;
  const double dgm = digamma((nu_old + (double)p) * 0.5) - log((nu_old + (double)p) * 0.5);
  const double deltanu = (nuhigh - nulow) / (double)Nd;
  const double sumq = -logsumw - dgm;
  int best = 0;
  double bestv = 0.0;
  for (int ci = 0; ci < Nd; ci++) {
    const double thisnu = nulow + (double)ci * deltanu;
    double q = -digamma(thisnu * 0.5) + log(thisnu * 0.5);
    q += -sumq + 1.0;
    if (ci == 0 || fabs(q) < bestv) {
      bestv = fabs(q);
      best = ci;
    }
  }
  return nulow + (double)best * deltanu;
}

// fns_fupdate_weights (rtr_solve_robust.c:296-383): row weights at x with the current nu (they stay
// on the device, folded into the tensors), then the new nu from mean(log w - w)
// The reference adds up its threads' partial sums of log w - w BEFORE joining the threads
// (rtr_solve_robust.c:361-370): the sums it reads
// depend on thread timing (observed: mostly still zero, i.e. mean(log w - w) = 0 and nu at the top
// of its grid; complete on small chunks).  nu_joined = true (default) is what the code was meant to
// do (sum after the join) and is pinned against a build of the reference with serialised threads
// (oracle/ref_shim_rtr_serial.c); false reproduces the "all sums still zero" outcome.
template <class EV>
static double update_weights(Ops<EV> &E, const double *x, double nu0, double nulow, double nuhigh,
                             bool keep, bool nu_joined) {
  double sumlogw = E.ev.weights_at(x, nu0, keep);
  if (!nu_joined) sumlogw = 0.0;
  const double nu1 = update_nu_aecm(sumlogw, nulow, nuhigh, 2, nu0);
  if (nu1 < nulow) return nulow;
  if (nu1 > nuhigh) return nuhigh;
  return nu1;
}

// Sartenaer's initial trust-region radius (itrr, rtr_solve_robust.c:1308-1437); moves x
template <class EV>
static double itrr(Ops<EV> &E, double *x, double *eta, double *Heta, double *s, double *x_prop) {
  const int n = E.n8;
  double delta_0 = 1.0, delta_m = 0.0, sigma = 0.0, delta = 0.0;
  double f0;
  E.fgrad(x, eta, true, &f0);
  const double eta_nrm = rtr_nrm2(n, eta);
  for (int i = 0; i < n; i++) eta[i] *= 1.0 / eta_nrm;
  for (int i = 0; i < n; i++) s[i] = eta[i] * delta_0;
  E.fhess(x, s, Heta);
  const double gamma_1 = 0.0625, gamma_2 = 5.0, gamma_3 = 0.5, gamma_4 = 2.0;
  const double mu_0 = 0.5, mu_1 = 0.5, mu_2 = 0.35, teta = 0.25;
  const int MK = 4;
  for (int m = 0; m < MK; m++) {
    for (int i = 0; i < n; i++) x_prop[i] = x[i] - s[i];
    const double mk = f0 - rtr_g(n, eta, s) - 0.5 * rtr_g(n, Heta, s);
    const double fk = E.f(x_prop);
    double rho;
    if (f0 == mk) rho = 1e9;
    else rho = (f0 - fk) / (f0 - mk);
    const double rho1 = fabs(rho - 1.0);
    if (rho1 < mu_0) delta_m = std::max(delta_m, delta_0);
    if ((f0 - fk) > delta) {
      delta = f0 - fk;
      sigma = delta_0;
    }
    double beta_i = 0.0;
    {
      const double g0_s = rtr_g(n, eta, s);
      const double b1 = (teta * (f0 - g0_s) + (1.0 - teta) * mk - fk);
      const double beta_1 = (b1 == 0.0 ? 1e9 : -teta * g0_s / b1);
      const double b2 = (-teta * (f0 - g0_s) + (1.0 + teta) * mk - fk);
      const double beta_2 = (b2 == 0.0 ? 1e9 : teta * g0_s / b2);
      const double minbeta = std::min(beta_1, beta_2), maxbeta = std::max(beta_1, beta_2);
      if (rho1 > mu_1) {
        if (minbeta > 1.0) beta_i = gamma_3;
        else if ((maxbeta < gamma_1) || (minbeta < gamma_1 && maxbeta >= 1.0)) beta_i = gamma_1;
        else if ((beta_1 >= gamma_1 && beta_1 < 1.0) && (beta_2 < gamma_1 || beta_2 >= 1.0)) beta_i = beta_1;
        else if ((beta_2 >= gamma_1 && beta_2 < 1.0) && (beta_1 < gamma_1 || beta_1 >= 1.0)) beta_i = beta_2;
        else beta_i = maxbeta;
      } else if (rho1 <= mu_2) {
        if (maxbeta < 1.0) beta_i = gamma_4;
        else if (maxbeta > gamma_2) beta_i = gamma_2;
        else if ((beta_1 >= 1.0 && beta_1 <= gamma_2) && beta_2 < 1.0) beta_i = beta_1;
        else if ((beta_2 >= 1.0 && beta_2 <= gamma_2) && beta_1 < 1.0) beta_i = beta_2;
        else beta_i = maxbeta;
      } else {
        if (maxbeta < gamma_3) beta_i = gamma_3;
        else if (maxbeta > gamma_4) beta_i = gamma_4;
        else beta_i = maxbeta;
      }
      delta_0 = delta_0 / beta_i;
    }
    for (int i = 0; i < n; i++) s[i] = eta[i] * delta_0;
  }
  if (delta > 0.0) rtr_axpy(n, eta, -sigma, x);
  return delta_m > 0.0 ? delta_m : delta_0;
}


// ------------------------------------------------------------------------------------------------
// one (cluster, chunk) solve.  kind: 4 RSD+RTR, 5 robust RTR, 6 robust NSD; aug != null (kind 5 only):
// rtr_solve_nocuda_robust_admm (rtr_solve_robust_admm.c:1424-1899, the same flow as kind 5 on the
// consensus-augmented cost in Euclidean space).  x: in/out (kept only if
// the cost went down).  robust_nu: in/out (lmdata.robust_nu persists from visit to visit,
// lmfit.c:938-957).  info[0] / info[1]: initial / final cost.
// ------------------------------------------------------------------------------------------------
template <class EV>
void solve_chunk(EV &ev, int kind, double *xio, int itmax_a, int itmax_b, double nulow,
                 double nuhigh, double *robust_nu, double *info, bool nu_joined = true,
                 const Admm *aug = nullptr) {
  Ops<EV> E(ev, aug);
  const int n8 = E.n8;
  std::vector<double> x(xio, xio + n8), fgradx(n8, 0.0), eta(n8, 0.0), Heta(n8, 0.0),
      x_prop(n8, 0.0);
  ev.unit_weights();
  E.count();
  double fx = E.f(x.data());
  const double fx0 = fx;
  if (kind == 4) {
    // RSD warm-up (rtr_solve.c:1318-1330)
    for (int ci = 0; ci < itmax_a; ci++) {
      const int rsdstat = armijostep(E, x.data(), eta.data(), fgradx.data(), x_prop.data(), &fx);
      for (int i = 0; i < n8; i++) x_prop[i] = x[i] + eta[i];
      if (!rsdstat) x = x_prop;
      else break;
    }
    const double Delta_bar = std::min(fx, 0.01);
    const double Delta0 = Delta_bar * 0.125;
    const double rho_regularization = fx * 1e-6;
    info[0] = fx;
    tr_loop(E, x.data(), fx, fgradx.data(), eta.data(), Heta.data(), x_prop.data(), itmax_b,
            Delta_bar, Delta0, rho_regularization);
    info[1] = fx;
  } else if (kind == 5) {
    const double Delta_new =
        itrr(E, x.data(), eta.data(), Heta.data(), fgradx.data(), x_prop.data());
    const double Delta0 = std::min(Delta_new, 0.01);
    const double Delta_bar = Delta0 * 8.0;
    const double rho_regularization = fx * 1e-6;
    *robust_nu = update_weights(E, x.data(), *robust_nu, nulow, nuhigh, true, nu_joined);
    info[0] = fx;
    tr_loop(E, x.data(), fx, fgradx.data(), eta.data(), Heta.data(), x_prop.data(), itmax_b,
            Delta_bar, Delta0, rho_regularization);
    info[1] = fx;
    *robust_nu = update_weights(E, x.data(), *robust_nu, nulow, nuhigh, false, nu_joined);
  } else {
    // Nesterov's accelerated steepest descent with unit weights (rtr_solve_robust.c:1990-2080)
    std::vector<double> z(n8), z_prop(n8);
    E.fgrad(x.data(), fgradx.data(), true);
    E.fhess(x.data(), x.data(), z.data());
    const double hess_nrm = rtr_nrm2(n8, z.data());
    double t = 1.0 / hess_nrm;
    if (t < 1e-6) t = 1e-6;
    z = x;
    double theta = 1.0;
    const double ALPHA = 1.01, BETA = 0.5;
    for (int it = 0; it < itmax_a; it++) {
      x_prop = x;
      z_prop = z;
      for (int i = 0; i < n8; i++) x[i] = z[i] - t * fgradx[i];
      const double grad_nrm = rtr_nrm2(n8, fgradx.data());
      const double x_nrm = rtr_nrm2(n8, x.data());
      if (grad_nrm * t / std::max(1.0, x_nrm) < 1e-6) break;
      theta = 2.0 / (1.0 + sqrt(1.0 + 4.0 / (theta * theta)));
      for (int i = 0; i < n8; i++) z[i] = (2.0 - theta) * x[i] - (1.0 - theta) * x_prop[i];
      eta = fgradx;
      E.fgrad(z.data(), fgradx.data(), true);
      for (int i = 0; i < n8; i++) {
        z_prop[i] -= z[i];
        eta[i] -= fgradx[i];
      }
      const double ydiffnrm = rtr_nrm2(n8, z_prop.data());
      double dot = 0.0;
      for (int i = 0; i < n8; i++) dot += z_prop[i] * eta[i];
      if (isnan(dot) || isinf(dot)) break;
      const double t_hat = 0.5 * (ydiffnrm * ydiffnrm) / fabs(dot);
      t = std::min(ALPHA * t, std::max(BETA * t, t_hat));
    }
    fx = E.f(x.data());
    info[1] = fx;  // info[0] is left as the previous visit set it (rtr_solve_robust.c:2086)
    *robust_nu = update_weights(E, x.data(), *robust_nu, nulow, nuhigh, false, nu_joined);
  }
  // the solution is kept only if the cost went down (rtr_solve.c:1583, rtr_solve_robust.c:1841,2098)
  if (fx0 > fx) memcpy(xio, x.data(), sizeof(double) * n8);
}

}  // namespace rtr
