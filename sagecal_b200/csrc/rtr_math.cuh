// Per-row and per-baseline arithmetic of the RTR-family kernels (kernels_rtr.cu), host-callable so
// that oracle/rtr_tensor_check.cu can run exactly this code on the CPU against the reference
// (test infrastructure; the product only runs it on the device).  See kernels_rtr.cu for the
// derivation.
#pragma once
#include "internal.cuh"

// the header is also compiled as host code: unroll pragmas only where they mean something
#ifdef __CUDA_ARCH__
#define RTR_UNROLL _Pragma("unroll")
#else
#define RTR_UNROLL
#endif

// accumulators of one baseline: upper triangle of T (diagonal + 6 complex), D, scalars
struct RtrAcc {
  double td[4];
  double2 to[6];
  double2 Dm[16];
  double c0, slw, cnt;
};
__host__ __device__ __forceinline__ void rtr_acc_zero(RtrAcc &A) {
RTR_UNROLL
  for (int i = 0; i < 4; i++) A.td[i] = 0.0;
RTR_UNROLL
  for (int i = 0; i < 6; i++) A.to[i] = make_double2(0, 0);
RTR_UNROLL
  for (int i = 0; i < 16; i++) A.Dm[i] = make_double2(0, 0);
  A.c0 = A.slw = A.cnt = 0.0;
}
// one unflagged row.  weighted: w = (nu+2)/(nu + max_c |d_c - (Gp C Gq^H)_c|^2)
// (threadfn_fns_fupdate_weights, rtr_solve_robust.c:210-262), slw += log w - w (:271-287)
__host__ __device__ __forceinline__ void rtr_acc_row(RtrAcc &A, const double2 *C, const double2 *dd,
                                                     bool weighted, const double2 *Gp,
                                                     const double2 *Gq, double nu, bool tensors) {
  double w = 1.0;
  if (weighted) {
    double2 T1[4], V[4];
    mat_ab(Gp, C, T1);
    mat_abh(T1, Gq, V);
    double mx = 0.0;
RTR_UNROLL
    for (int c = 0; c < 4; c++) {
      const double er = dd[c].x - V[c].x, ei = dd[c].y - V[c].y;
      const double e2 = er * er + ei * ei;
      mx = e2 > mx ? e2 : mx;
    }
    w = (nu + 2.0) / (nu + mx);
    A.slw += log(w) - w;
  }
  A.cnt += 1.0;
  if (!tensors) return;
  double2 wC[4];
RTR_UNROLL
  for (int c = 0; c < 4; c++) wC[c] = make_double2(w * C[c].x, w * C[c].y);
  // T[(i),(k)] = sum w C_i conj(C_k), i <= k
  int o = 0;
RTR_UNROLL
  for (int i = 0; i < 4; i++) {
    A.td[i] = fma(wC[i].x, C[i].x, fma(wC[i].y, C[i].y, A.td[i]));
RTR_UNROLL
    for (int k = i + 1; k < 4; k++) cfmac(A.to[o++], wC[i], C[k]);
  }
RTR_UNROLL
  for (int ab = 0; ab < 4; ab++) {
    A.c0 = fma(w * dd[ab].x, dd[ab].x, fma(w * dd[ab].y, dd[ab].y, A.c0));
RTR_UNROLL
    for (int ij = 0; ij < 4; ij++) cfmac(A.Dm[4 * ab + ij], dd[ab], wC[ij]);
  }
}
// full 4x4 T (row-major) and D from the accumulators
__host__ __device__ __forceinline__ void rtr_acc_expand(const RtrAcc &A, double2 *T, double2 *D) {
  int o = 0;
RTR_UNROLL
  for (int i = 0; i < 4; i++) {
    T[4 * i + i] = make_double2(A.td[i], 0.0);
RTR_UNROLL
    for (int k = i + 1; k < 4; k++) {
      T[4 * i + k] = A.to[o];
      T[4 * k + i] = make_double2(A.to[o].x, -A.to[o].y);
      o++;
    }
  }
RTR_UNROLL
  for (int i = 0; i < 16; i++) D[i] = A.Dm[i];
}

// M[(ab),(mj)] = sum_{i,j'} A1_ai conj(A2_bj') T[(ij'),(mj)]   (acc: M += ...)
__host__ __device__ __forceinline__ void rtr_M(const double2 *A1, const double2 *A2, const double2 *T,
                                      double2 *M, bool acc) {
  // Z[(a j'),(mj)] = sum_i A1_ai T[(i j'),(mj)]
  double2 Z[16];
RTR_UNROLL
  for (int a = 0; a < 2; a++)
RTR_UNROLL
    for (int jp = 0; jp < 2; jp++)
RTR_UNROLL
      for (int mj = 0; mj < 4; mj++)
        Z[4 * (2 * a + jp) + mj] =
            cdot2(A1[2 * a + 0], T[4 * (0 * 2 + jp) + mj], A1[2 * a + 1], T[4 * (1 * 2 + jp) + mj]);
RTR_UNROLL
  for (int a = 0; a < 2; a++)
RTR_UNROLL
    for (int b = 0; b < 2; b++)
RTR_UNROLL
      for (int mj = 0; mj < 4; mj++) {
        double2 v = acc ? M[4 * (2 * a + b) + mj] : make_double2(0, 0);
        cfmacl(v, A2[2 * b + 0], Z[4 * (2 * a + 0) + mj]);
        cfmacl(v, A2[2 * b + 1], Z[4 * (2 * a + 1) + mj]);
        M[4 * (2 * a + b) + mj] = v;
      }
}

// one baseline seen from station s (sp: s is the baseline's p).  T, W: tensors T and D of the baseline
// (W is overwritten), c0: its sum w |d|^2.  Adds the station sums of the gradient (hess == false) or of
// the Hessian-vector product to acc[4] (if want_vec) and, from the p end only, the cost to *cost.
__host__ __device__ __forceinline__ void rtr_eval_baseline(bool sp, const double2 *Gs,
                                                           const double2 *Go, const double2 *Es,
                                                           const double2 *Eo, const double2 *T,
                                                           double2 *W, double c0, bool hess,
                                                           bool want_cost, bool want_vec,
                                                           double2 *acc, double *cost) {
  const double2 *Gp = sp ? Gs : Go, *Gq = sp ? Go : Gs;
  // Wres = D - M(Gp,Gq); the cost needs D + Wres
  double2 Mg[16];
  rtr_M(Gp, Gq, T, Mg, false);
  if (want_cost && sp) {
    double cs = c0;
RTR_UNROLL
    for (int aa = 0; aa < 2; aa++)
RTR_UNROLL
      for (int bb = 0; bb < 2; bb++)
RTR_UNROLL
        for (int i = 0; i < 2; i++)
RTR_UNROLL
          for (int j = 0; j < 2; j++) {
            // conj(K) = conj(Gp_ai) Gq_bj ;  (D + Wres) = 2D - M
            const double2 K = cmulc(Gp[2 * aa + i], Gq[2 * bb + j]);  // Gp_ai conj(Gq_bj)
            const int e = 4 * (2 * aa + bb) + (2 * i + j);
            const double vr = 2.0 * W[e].x - Mg[e].x, vi = 2.0 * W[e].y - Mg[e].y;
            cs -= K.x * vr + K.y * vi;  // Re(conj(K) v)
          }
    *cost += cs;
  }
  if (!want_vec) return;
RTR_UNROLL
  for (int i = 0; i < 16; i++) W[i] = csub(W[i], Mg[i]);
  if (!hess) {
    if (sp) {
      // grad_p[a,m] += sum_bj Gq_bj Wres[(ab),(mj)]
RTR_UNROLL
      for (int aa = 0; aa < 2; aa++)
RTR_UNROLL
        for (int m = 0; m < 2; m++)
RTR_UNROLL
          for (int bb = 0; bb < 2; bb++)
RTR_UNROLL
            for (int j = 0; j < 2; j++)
              cfma(acc[2 * aa + m], Gq[2 * bb + j], W[4 * (2 * aa + bb) + (2 * m + j)]);
    } else {
      // grad_q[b,m] += sum_ai Gp_ai conj(Wres[(ab),(im)])
RTR_UNROLL
      for (int bb = 0; bb < 2; bb++)
RTR_UNROLL
        for (int m = 0; m < 2; m++)
RTR_UNROLL
          for (int aa = 0; aa < 2; aa++)
RTR_UNROLL
            for (int i = 0; i < 2; i++)
              cfmac(acc[2 * bb + m], Gp[2 * aa + i], W[4 * (2 * aa + bb) + (2 * i + m)]);
    }
  } else {
    const double2 *Ep = sp ? Es : Eo, *Eq = sp ? Eo : Es;
    double2 W1[16];
    rtr_M(Gp, Eq, T, W1, false);
    rtr_M(Ep, Gq, T, W1, true);
    if (sp) {
RTR_UNROLL
      for (int aa = 0; aa < 2; aa++)
RTR_UNROLL
        for (int m = 0; m < 2; m++)
RTR_UNROLL
          for (int bb = 0; bb < 2; bb++)
RTR_UNROLL
            for (int j = 0; j < 2; j++) {
              const int e = 4 * (2 * aa + bb) + (2 * m + j);
              cfma(acc[2 * aa + m], Eq[2 * bb + j], W[e]);
              double2 ng = make_double2(-Gq[2 * bb + j].x, -Gq[2 * bb + j].y);
              cfma(acc[2 * aa + m], ng, W1[e]);
            }
    } else {
RTR_UNROLL
      for (int bb = 0; bb < 2; bb++)
RTR_UNROLL
        for (int m = 0; m < 2; m++)
RTR_UNROLL
          for (int aa = 0; aa < 2; aa++)
RTR_UNROLL
            for (int i = 0; i < 2; i++) {
              const int e = 4 * (2 * aa + bb) + (2 * i + m);
              cfmac(acc[2 * bb + m], Ep[2 * aa + i], W[e]);
              double2 ng = make_double2(-Gp[2 * aa + i].x, -Gp[2 * aa + i].y);
              cfmac(acc[2 * bb + m], ng, W1[e]);
            }
    }
  }
}

// ---- the same evaluation spread over 16 lanes per baseline end (k_rtr_eval) ---------------------------
// lane l of a baseline end: a = l>>3 & 1, b = l>>2 & 1, m = l>>1 & 1, j = l & 1 owns the tensor entry
// [(ab),(mj)].  Tc[2i+j'] = T[(i j'),(mj)] (the lane's column of T), Dv = D[(ab),(mj)].
//   *term: the lane's addend of the station sums: seen from p (sp) it belongs to entry (a,m) of the
//          station's 2x2 block and is summed over the 4 lanes (b,j); seen from q it belongs to entry
//          (b,j) and is summed over the 4 lanes (a,m) -- sums are linear, so the caller adds the lanes'
//          terms over all ends first and folds the 4 lanes once at the end
//   *cost: the lane's share of sum_t w |d - Gp C Gq^H|^2 without c0 (p end only)
// rows: Gpa = row a of Gp, Gqb = row b of Gq, Epa / Eqb likewise (two entries each; the lane-dependent
// picks below are selects, not indexed register arrays)
__host__ __device__ __forceinline__ double2 rtr_lane_M(const double2 *A1row, const double2 *A2row,
                                                       const double2 *Tc) {
  // sum_{i,j'} A1_ai conj(A2_bj') T[(ij'),(mj)]
  const double2 z0 = cdot2(A1row[0], Tc[0], A1row[1], Tc[2]);  // j' = 0
  const double2 z1 = cdot2(A1row[0], Tc[1], A1row[1], Tc[3]);  // j' = 1
  double2 v = make_double2(0.0, 0.0);
  cfmacl(v, A2row[0], z0);
  cfmacl(v, A2row[1], z1);
  return v;
}
__host__ __device__ __forceinline__ void rtr_lane_terms(int lane, bool sp, const double2 *Gpa,
                                                        const double2 *Gqb, const double2 *Epa,
                                                        const double2 *Eqb, const double2 *Tc,
                                                        double2 Dv, bool hess, bool want_cost,
                                                        bool want_vec, double2 *term,
                                                        double *cost) {
  const int m = (lane >> 1) & 1, j = lane & 1;
  const double2 Gpm = m ? Gpa[1] : Gpa[0], Gqj = j ? Gqb[1] : Gqb[0];
  const double2 Mg = rtr_lane_M(Gpa, Gqb, Tc);
  if (want_cost && sp) {
    const double2 K = cmulc(Gpm, Gqj);  // Gp_ai conj(Gq_bj), (ij) = (mj)
    const double vr = 2.0 * Dv.x - Mg.x, vi = 2.0 * Dv.y - Mg.y;
    *cost -= K.x * vr + K.y * vi;
  }
  if (!want_vec) return;
  const double2 W = csub(Dv, Mg);  // Wres[(ab),(mj)]
  double2 t = make_double2(0.0, 0.0);
  if (!hess) {
    if (sp) cfma(t, Gqj, W);         // grad_p[a,m] += Gq_bj Wres[(ab),(mj)]
    else cfmac(t, Gpm, W);           // grad_q[b,m'] += Gp_ai conj(Wres[(ab),(i m')]), i = m bit
  } else {
    const double2 W1 = cadd(rtr_lane_M(Gpa, Eqb, Tc), rtr_lane_M(Epa, Gqb, Tc));
    if (sp) {
      const double2 Eqj = j ? Eqb[1] : Eqb[0];
      cfma(t, Eqj, W);
      cfma(t, make_double2(-Gqj.x, -Gqj.y), W1);
    } else {
      const double2 Epm = m ? Epa[1] : Epa[0];
      cfmac(t, Epm, W);
      cfmac(t, make_double2(-Gpm.x, -Gpm.y), W1);
    }
  }
  *term = t;
}
// entry e = 2 row + col of the station's 2x2 block from the 16 lane sums of the p ends (tP) and of the
// q ends (tQ): p ends fold the lanes (b,j), q ends the lanes (a,m)
__host__ __device__ __forceinline__ double2 rtr_lane_fold(const double2 *tP, const double2 *tQ,
                                                          int e) {
  const int r = e >> 1, c = e & 1;
  double2 v = make_double2(0.0, 0.0);
  for (int x = 0; x < 2; x++)
    for (int y = 0; y < 2; y++) {
      v = cadd(v, tP[r * 8 + x * 4 + c * 2 + y]);   // a = r, m = c, over b = x, j = y
      v = cadd(v, tQ[x * 8 + r * 4 + y * 2 + c]);   // b = r, j = c, over a = x, m = y
    }
  return v;
}
