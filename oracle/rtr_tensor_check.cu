// TEST INFRASTRUCTURE ONLY (oracle).  Runs the arithmetic of the product's RTR-family kernels
// (sagecal_b200/csrc/rtr_math.cuh: the per-row condensation into per-baseline tensors and the
// per-baseline cost / gradient / Hessian evaluation) on the CPU, under the control flow of
// rtr_algo.h, so that the tensor formulation is pinned against the compiled reference without a
// GPU (tests/test_oracle_rtr.py).  Compiled by nvcc as host code; nothing is launched.
#include <math.h>
#include <vector>

#include "../sagecal_b200/csrc/rtr_algo.h"
#include "../sagecal_b200/csrc/rtr_math.cuh"

namespace {
// mirror of orc_problem (dirac_oracle.h), the coherencies as doubles
struct OrcProblem {
  int N, Nbase, tilesz, M, Mt;
  const int *sta1, *sta2;
  const unsigned char *flag;
  const int *nchunk, *chunk0, *chunk_off;
  const double *coh;  // [row][M][4] complex
};

struct TensorEval {
  const OrcProblem *P;
  int k, t0, ntiles, N;
  const double *y;
  std::vector<double2> T, D;   // [Nbase][16]
  std::vector<double> c0, cnt;

  static void jones(const double *x, int s, double2 *J) {
    for (int i = 0; i < 4; i++) J[i] = make_double2(x[8 * s + 2 * i], x[8 * s + 2 * i + 1]);
  }
  // what k_rtr_stats does, baseline by baseline
  double condense(const double *xw, double nu, bool tensors) {
    const int Nbase = P->Nbase;
    double slw = 0.0;
    if (tensors) { T.assign((size_t)Nbase * 16, make_double2(0, 0)); D = T; c0.assign(Nbase, 0.0); }
    cnt.assign(Nbase, 0.0);
    for (int b = 0; b < Nbase; b++) {
      RtrAcc A;
      rtr_acc_zero(A);
      double2 Gp[4], Gq[4];
      for (int t = 0; t < ntiles; t++) {
        const long r = (long)(t0 + t) * Nbase + b, ry = (long)t * Nbase + b;
        if (P->flag[r]) continue;
        if (xw) { jones(xw, P->sta1[r], Gp); jones(xw, P->sta2[r], Gq); }
        double2 C[4], dd[4];
        for (int c = 0; c < 4; c++) {
          const double *cc = P->coh + 2 * (4 * (size_t)P->M * r + 4 * k + c);
          C[c] = make_double2(cc[0], cc[1]);
          dd[c] = make_double2(y[8 * ry + 2 * c], y[8 * ry + 2 * c + 1]);
        }
        rtr_acc_row(A, C, dd, xw != nullptr, Gp, Gq, nu, tensors);
      }
      slw += A.slw;
      cnt[b] = A.cnt;
      if (tensors) {
        rtr_acc_expand(A, &T[(size_t)b * 16], &D[(size_t)b * 16]);
        c0[b] = A.c0;
      }
    }
    return slw;
  }
  // what k_rtr_eval does: station by station, every baseline end spread over 16 lanes, the lanes'
  // terms summed over the ends, folded once per station
  void raw(const double *x, const double *eta, double *fcost, double *vec) {
    double cost = 0.0;
    for (int s = 0; s < N; s++) {
      double2 Gs[4], Es[4], tP[16], tQ[16];
      jones(x, s, Gs);
      if (eta) jones(eta, s, Es);
      for (int i = 0; i < 16; i++) tP[i] = tQ[i] = make_double2(0, 0);
      for (int o = 0; o < N; o++) {
        if (o == s) continue;
        const bool sp = s < o;
        const int p = sp ? s : o, q = sp ? o : s;
        const size_t b = (size_t)baseline_index(p, q, N);
        if (!vec && !(fcost && sp)) continue;
        double2 Go[4], Eo[4];
        jones(x, o, Go);
        if (eta) jones(eta, o, Eo);
        if (fcost && sp) cost += c0[b];
        for (int lane = 0; lane < 16; lane++) {
          const int mj = lane & 3, ab = lane >> 2;
          double2 Tc[4];
          for (int k = 0; k < 4; k++) Tc[k] = T[b * 16 + 4 * k + mj];
          double2 term = make_double2(0, 0);
          const int la = (lane >> 3) & 1, lb = (lane >> 2) & 1;
          const double2 *Gp = sp ? Gs : Go, *Gq = sp ? Go : Gs, *Ep = sp ? Es : Eo, *Eq = sp ? Eo : Es;
          rtr_lane_terms(lane, sp, Gp + 2 * la, Gq + 2 * lb, Ep + 2 * la, Eq + 2 * lb, Tc,
                         D[b * 16 + 4 * ab + mj], eta != nullptr, fcost != nullptr, vec != nullptr,
                         &term, &cost);
          if (sp) tP[lane] = cadd(tP[lane], term);
          else tQ[lane] = cadd(tQ[lane], term);
        }
      }
      if (vec)
        for (int e = 0; e < 4; e++) {
          const double2 v = rtr_lane_fold(tP, tQ, e);
          vec[8 * s + 2 * e] = v.x;
          vec[8 * s + 2 * e + 1] = v.y;
        }
    }
    if (fcost) *fcost = cost;
  }
  void counts(double *c) {
    for (int s = 0; s < N; s++) {
      double v = 0.0;
      for (int o = 0; o < N; o++)
        if (o != s) v += cnt[(size_t)baseline_index(s < o ? s : o, s < o ? o : s, N)];
      c[s] = v;
    }
  }
  void unit_weights() { condense(nullptr, 0.0, true); }
  double weights_at(const double *x, double nu, bool keep) {
    return condense(x, nu, keep) / (double)((long)ntiles * P->Nbase);
  }
};
}  // namespace

extern "C" void harness_rtr_solve_tensor(const void *P, int k, int t0, int ntiles, const double *y,
                                         int kind, double *x, int itmax_a, int itmax_b,
                                         double nulow, double nuhigh, double *robust_nu,
                                         double *info, int nu_joined, const double *Y,
                                         const double *BZ, double rho) {
  TensorEval E;
  E.P = (const OrcProblem *)P;
  E.k = k; E.t0 = t0; E.ntiles = ntiles; E.N = E.P->N; E.y = y;
  rtr::Admm aug = {Y, BZ, rho};
  rtr::solve_chunk(E, kind, x, itmax_a, itmax_b, nulow, nuhigh, robust_nu, info, nu_joined != 0,
                   Y ? &aug : nullptr);
}
