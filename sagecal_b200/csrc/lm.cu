// Per-cluster Levenberg-Marquardt on the device-resident problem.
//
// Control flow mirrors clevmar_der_single_nocuda (clmfit.c:219-529) and
// oslevmar_der_single_nocuda (clmfit.c:1281-1640) decision for decision; what differs is how the
// quantities are produced:
//   e, ||e||^2, J^T e   one streaming pass over (hidden data, coh_k)          k_cluster_pass
//   J^T J               assembled from the per-baseline Gram tensors           k_coh_gram/k_assemble
//   (J^T J + mu I) dp   cuSOLVER potrf/potrs | geqrf+ormqr+trsm | gesvd        (library, not HBM bound)
// The dense n x 8N Jacobian of the reference (7.2 GB per cluster at N=62, T=120) never exists.
#include <float.h>
#include <math.h>
#include <string.h>

#include "../../include/dirac_b200.h"
#include "problem.h"

#define CS_CHECK(call)                                                                     \
  do {                                                                                     \
    cusolverStatus_t s__ = (call);                                                         \
    if (s__ != CUSOLVER_STATUS_SUCCESS) {                                                  \
      fprintf(stderr, "dirac_b200: cuSOLVER error %d at %s:%d\n", (int)s__, __FILE__, __LINE__); \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)
#define CB_CHECK(call)                                                                     \
  do {                                                                                     \
    cublasStatus_t s__ = (call);                                                           \
    if (s__ != CUBLAS_STATUS_SUCCESS) {                                                    \
      fprintf(stderr, "dirac_b200: cuBLAS error %d at %s:%d\n", (int)s__, __FILE__, __LINE__); \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

template <typename T>
static T *dalloc(size_t n) {
  T *p = nullptr;
  DB_CHECK(cudaMalloc((void **)&p, n * sizeof(T) + 16));
  return p;
}

void db_lm_init(dirac_b200_problem *pr) {
  LMWork &w = pr->lm;
  if (w.ready) return;
  DevProblem &d = pr->d;
  const int n8 = 8 * d.N;
  w.n8 = n8;
  w.T = dalloc<double>((size_t)d.Mt * d.Nbase * 16);
  w.T_valid = (unsigned char *)calloc(d.Mt, 1);
  w.Tsub = dalloc<double>((size_t)d.Nbase * 16);
  w.JTJ0 = dalloc<double>((size_t)n8 * n8);
  w.JTJ = dalloc<double>((size_t)n8 * n8);
  w.JTe = dalloc<double>(n8);
  w.JTe_new = dalloc<double>(n8);
  w.Hst = dalloc<double>((size_t)4 * d.N);
  w.Dp = dalloc<double>(n8);
  w.pnew = dalloc<double>(n8);
  w.devinfo = dalloc<int>(4);
  w.tau = dalloc<double>(n8);
  w.svdS = w.svdU = w.svdVT = nullptr;
  DB_CHECK(cudaMallocHost((void **)&w.h_vec, sizeof(double) * (4 * n8 + 4 * d.N + 16)));
  CS_CHECK(cusolverDnCreate(&w.cs));
  CS_CHECK(cusolverDnSetStream(w.cs, d.stream));
  CB_CHECK(cublasCreate(&w.cb));
  CB_CHECK(cublasSetStream(w.cb, d.stream));
  int l1 = 0, l2 = 0, l3 = 0;
  CS_CHECK(cusolverDnDpotrf_bufferSize(w.cs, CUBLAS_FILL_MODE_LOWER, n8, w.JTJ, n8, &l1));
  CS_CHECK(cusolverDnDgeqrf_bufferSize(w.cs, n8, n8, w.JTJ, n8, &l2));
  CS_CHECK(cusolverDnDormqr_bufferSize(w.cs, CUBLAS_SIDE_LEFT, CUBLAS_OP_T, n8, 1, n8, w.JTJ, n8,
                                       w.tau, w.Dp, n8, &l3));
  w.lwork = l1 > l2 ? l1 : l2;
  if (l3 > w.lwork) w.lwork = l3;
  int l4 = 0;
  CS_CHECK(cusolverDnDgesvd_bufferSize(w.cs, n8, n8, &l4));
  if (l4 > w.lwork) w.lwork = l4;
  w.cswork = dalloc<double>((size_t)w.lwork);
  w.dbuf = dalloc<double2>((size_t)4 * d.R);
  w.ready = true;
}

void db_lm_free(dirac_b200_problem *pr) {
  LMWork &w = pr->lm;
  if (!w.ready) return;
  cudaFree(w.T); cudaFree(w.Tsub); cudaFree(w.JTJ0); cudaFree(w.JTJ); cudaFree(w.JTe);
  cudaFree(w.JTe_new); cudaFree(w.Hst); cudaFree(w.Dp); cudaFree(w.pnew); cudaFree(w.devinfo);
  cudaFree(w.tau); cudaFree(w.cswork); cudaFree(w.dbuf);
  if (w.svdS) { cudaFree(w.svdS); cudaFree(w.svdU); cudaFree(w.svdVT); }
  cudaFreeHost(w.h_vec);
  free(w.T_valid);
  cusolverDnDestroy(w.cs);
  cublasDestroy(w.cb);
  w.ready = false;
}

// timeslots per CTA slice of a per-cluster pass: enough slices to fill the GPU, long enough to
// amortise the per-slice station reduction
static int pick_tslice(const DevProblem &d, int nt) {
  int target_ctas = 148 * 2;
  int slices = (target_ctas + d.ntile - 1) / d.ntile;
  if (slices < 1) slices = 1;
  int ts = (nt + slices - 1) / slices;
  if (ts < 2) ts = 2;
  if (ts > nt) ts = nt;
  if (ts < 1) ts = 1;
  return ts;
}

// one streaming pass of cluster k over timeslots [t0,t1): see ClusterPassArgs for the modes
void db_cluster_pass(dirac_b200_problem *pr, int k, const double *pblk_dev, const double2 *in,
                     double2 *out, int mode, int write_out, double *jte_dev, int cost_slot, int t0,
                     int t1) {
  DevProblem &d = pr->d;
  if (t1 <= t0) {
    if (mode <= 1) DB_CHECK(cudaMemsetAsync(d.scal + cost_slot, 0, sizeof(double), d.stream));
    if (jte_dev) DB_CHECK(cudaMemsetAsync(jte_dev, 0, sizeof(double) * 8 * d.N, d.stream));
    return;
  }
  ClusterPassArgs a;
  a.coh_k = d.coh + (size_t)k * 4 * d.R;
  a.in = in; a.flag = d.flag; a.pblk = pblk_dev; a.tiles = d.tiles; a.out = out; a.jte = jte_dev;
  a.partials = pr->partials; a.cost = d.scal + cost_slot; a.counter = d.counters; a.R = d.R;
  a.N = d.N; a.Nbase = d.Nbase; a.t_begin = t0; a.t_end = t1; a.tslice = pick_tslice(d, t1 - t0);
  a.mode = mode; a.write_out = write_out;
  if (jte_dev && mode <= 1)
    DB_CHECK(cudaMemsetAsync(jte_dev, 0, sizeof(double) * 8 * d.N, d.stream));
  db_prof_begin(2, (double)(t1 - t0) * d.Nbase * (129.0 + (write_out ? 64.0 : 0.0)), d.stream);
  db_launch_cluster_pass(&a, d.ntile, d.stream);
  db_prof_end(d.stream);
  db_count_launch(1);
}

// Gram tensor of cluster k over timeslots t0, t0+step, ... < t1 into Tdst [Nbase][16]
static void gram(dirac_b200_problem *pr, int k, int t0, int t1, int step, double *Tdst) {
  DevProblem &d = pr->d;
  GramArgs a;
  a.coh = d.coh; a.flag = d.flag; a.tiles = d.tiles; a.T = Tdst; a.R = d.R; a.N = d.N;
  a.Nbase = d.Nbase; a.k0 = k; a.t_begin = t0; a.t_end = t1; a.t_step = step;
  db_prof_begin(3, (double)((t1 - t0 + step - 1) / step) * d.Nbase * 65.0 + 128.0 * d.Nbase, d.stream);
  db_launch_coh_gram(&a, d.ntile, 1, d.stream);
  db_prof_end(d.stream);
  db_count_launch(1);
}

static void assemble(dirac_b200_problem *pr, const double *T, const double *pblk_dev,
                     double *JTJ) {
  DevProblem &d = pr->d;
  LMWork &w = pr->lm;
  DB_CHECK(cudaMemsetAsync(w.Hst, 0, sizeof(double) * 4 * d.N, d.stream));
  AssembleArgs a;
  a.T = T; a.pblk = pblk_dev; a.JTJ = JTJ; a.Hst = w.Hst; a.tiles = d.tiles; a.N = d.N;
  a.Nbase = d.Nbase;
  db_prof_begin(4, 128.0 * d.Nbase + 8.0 * 64.0 * d.N * d.N, d.stream);
  db_launch_assemble(&a, d.ntile, d.stream);
  db_prof_end(d.stream);
  db_count_launch(2);
}

// chunk ck of cluster k covers timeslots [t0,t1)  (lmfit.c:893-905)
static void chunk_range(const DevProblem &d, int k, int ck, int *t0, int *t1) {
  int nchunk = d.h_clus[k].nchunk;
  int tilechunk = (d.tilesz + nchunk - 1) / nchunk;
  int a = ck * tilechunk;
  int b = a + tilechunk;
  if (a > d.tilesz) a = d.tilesz;
  if (b > d.tilesz) b = d.tilesz;
  *t0 = a;
  *t1 = b;
}

// ------------------------------------------------------------------------------------------------
// damped solve (J^T J + mu I) dp = J^T e on the device.  returns 1 if solved.
// linsolv: 0 Cholesky (dpotrf/dpotrs, clmfit.c:373-395), 1 QR (dgels, :396-409),
//          2 SVD with singular-value cut at eps1 (:410-436)
// ------------------------------------------------------------------------------------------------
static int damped_solve(dirac_b200_problem *pr, double mu, int linsolv, double eps1) {
  DevProblem &d = pr->d;
  LMWork &w = pr->lm;
  const int n = w.n8;
  db_launch_copy_add_diag(w.JTJ0, w.JTJ, n, mu, d.stream);
  db_count_launch(1);
  DB_CHECK(cudaMemcpyAsync(w.Dp, w.JTe, sizeof(double) * n, cudaMemcpyDeviceToDevice, d.stream));
  int *hinfo = (int *)(w.h_vec + 4 * n + 4 * d.N);
  hinfo[0] = hinfo[1] = 0;
  db_prof_begin(5, 0.0, d.stream);
  if (linsolv == 0) {
    CS_CHECK(cusolverDnDpotrf(w.cs, CUBLAS_FILL_MODE_LOWER, n, w.JTJ, n, w.cswork, w.lwork,
                              w.devinfo));
    CS_CHECK(cusolverDnDpotrs(w.cs, CUBLAS_FILL_MODE_LOWER, n, 1, w.JTJ, n, w.Dp, n,
                              w.devinfo + 1));
    db_count_launch(2);
    DB_CHECK(cudaMemcpyAsync(hinfo, w.devinfo, 2 * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
  } else if (linsolv == 1) {
    // A = QR ; dp = R^-1 Q^T b   (A symmetric: row/column-major views coincide)
    CS_CHECK(cusolverDnDgeqrf(w.cs, n, n, w.JTJ, n, w.tau, w.cswork, w.lwork, w.devinfo));
    CS_CHECK(cusolverDnDormqr(w.cs, CUBLAS_SIDE_LEFT, CUBLAS_OP_T, n, 1, n, w.JTJ, n, w.tau, w.Dp,
                              n, w.cswork, w.lwork, w.devinfo + 1));
    const double one = 1.0;
    CB_CHECK(cublasDtrsm(w.cb, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N,
                         CUBLAS_DIAG_NON_UNIT, n, 1, &one, w.JTJ, n, w.Dp, n));
    db_count_launch(3);
    DB_CHECK(cudaMemcpyAsync(hinfo, w.devinfo, 2 * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
  } else {
    if (!w.svdS) {
      w.svdS = dalloc<double>(n);
      w.svdU = dalloc<double>((size_t)n * n);
      w.svdVT = dalloc<double>((size_t)n * n);
    }
    CS_CHECK(cusolverDnDgesvd(w.cs, 'A', 'A', n, n, w.JTJ, n, w.svdS, w.svdU, n, w.svdVT, n,
                              w.cswork, w.lwork, nullptr, w.devinfo));
    db_count_launch(1);
    // dp = V diag(1/s | s>eps1) U^T b, small: finish on the host vectors
    double *hS = (double *)malloc(sizeof(double) * n);
    double *hb = (double *)malloc(sizeof(double) * n);
    const double one = 1.0, zero = 0.0;
    CB_CHECK(cublasDgemv(w.cb, CUBLAS_OP_T, n, n, &one, w.svdU, n, w.JTe, 1, &zero, w.Dp, 1));
    DB_CHECK(cudaMemcpyAsync(hS, w.svdS, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream));
    DB_CHECK(cudaMemcpyAsync(hb, w.Dp, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream));
    DB_CHECK(cudaStreamSynchronize(d.stream));
    for (int i = 0; i < n; i++) hb[i] = (hS[i] > eps1) ? hb[i] / hS[i] : 0.0;
    DB_CHECK(cudaMemcpyAsync(w.pnew, hb, sizeof(double) * n, cudaMemcpyHostToDevice, d.stream));
    CB_CHECK(cublasDgemv(w.cb, CUBLAS_OP_T, n, n, &one, w.svdVT, n, w.pnew, 1, &zero, w.Dp, 1));
    db_prof_end(d.stream);
    DB_CHECK(cudaStreamSynchronize(d.stream));
    free(hS);
    free(hb);
    db_count_launch(2);
    return 1;
  }
  db_prof_end(d.stream);
  // the step itself comes back with the status
  DB_CHECK(cudaMemcpyAsync(w.h_vec + 2 * n, w.Dp, sizeof(double) * n, cudaMemcpyDeviceToHost,
                           d.stream));
  DB_CHECK(cudaStreamSynchronize(d.stream));
  return (hinfo[0] == 0 && hinfo[1] == 0) ? 1 : 0;
}

static double nrm2sq(const double *v, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += v[i] * v[i];
  return s;
}

// ------------------------------------------------------------------------------------------------
// LM on chunk ck of cluster k.  `r` holds the residual of the full model over the whole interval;
// on return the chunk's rows of `r` are the residual with the updated Jones.  pblk_dev points at
// the 8N parameters inside the device copy of pp (updated in place).
//   os != 0 : ordered-subsets variant (clmfit.c:1074): J^T J and J^T e of ONE of Nsubsets time
//             subsets per iteration.
// info[0] = ||e||^2 at entry, info[1] = ||e||^2 at exit (lmfit.c:963-964 uses exactly these).
// ------------------------------------------------------------------------------------------------
void db_lm_chunk(dirac_b200_problem *pr, int k, int ck, double *pblk_dev, double2 *r, int itmax,
                 const double *opts, int linsolv, int os, int randomize, double *info) {
  DevProblem &d = pr->d;
  db_lm_init(pr);
  LMWork &w = pr->lm;
  const int n = w.n8;
  int t0, t1;
  chunk_range(d, k, ck, &t0, &t1);
  const int ntiles = t1 - t0;
  const double tau = opts ? opts[0] : 1e-3;
  const double eps1 = opts ? opts[1] : 1e-17;
  const double eps2 = opts ? opts[2] : 1e-17;
  const double eps2_sq = eps2 * eps2;
  const double eps3 = opts ? opts[3] : 1e-17;

  double *hp = w.h_vec;            // current p
  double *hjte = w.h_vec + n;      // J^T e
  double *hDp = w.h_vec + 2 * n;   // step
  double *hpnew = w.h_vec + 3 * n; // trial p
  double *hH = w.h_vec + 4 * n;    // station sums [N][4]

  // hidden data d = r + f(p_old); e = d - f(p_old); ||e||^2; J^T e  (lmfit.c:890-891 fused with
  // the first func/jacf evaluation, clmfit.c:241-252)
  db_cluster_pass(pr, k, pblk_dev, r, w.dbuf, 0, 1, os ? nullptr : w.JTe, 1, t0, t1);
  DB_CHECK(cudaMemcpyAsync(hp, pblk_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream));
  if (!os)
    DB_CHECK(cudaMemcpyAsync(hjte, w.JTe, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream));
  double p_eL2 = db_read_scalar(pr, 1);  // synchronises
  const double init_p_eL2 = p_eL2;
  int stop = 0;
  if (!isfinite(p_eL2)) stop = 7;
  int nu = 2, nu2;
  double mu = 0.0, Dp_L2 = DBL_MAX, jacTe_inf = 0.0;

  // ordered subsets (clmfit.c:1313-1356)
  int Nsubsets = 10;
  if (ntiles < Nsubsets) Nsubsets = ntiles;
  const int max_os_iter = os ? (int)ceil(0.1 * (double)Nsubsets) : 1;
  const int Ntper = os && Nsubsets > 0 ? (ntiles + Nsubsets - 1) / Nsubsets : ntiles;

  // Gram tensor of this chunk (time-invariant part of J^T J), built once per solve interval
  const int tix = d.h_clus[k].chunk0 + ck;
  double *Tfull = w.T + (size_t)tix * d.Nbase * 16;
  if (!os && !w.T_valid[tix] && ntiles > 0) {
    gram(pr, k, t0, t1, 1, Tfull);
    w.T_valid[tix] = 1;
  }

  int kiter;
  for (kiter = 0; kiter < itmax && !stop; ++kiter) {
    if (p_eL2 <= eps3) {
      stop = 6;
      break;
    }
    for (int ositer = 0; ositer < max_os_iter; ositer++) {
      const double *Tuse = Tfull;
      if (os) {
        int l;
        if (randomize) {
          l = rand() % Nsubsets;  // FIXME: the reference draws a permutation (clmfit.c:1372)
        } else {
          l = (kiter + ositer) % Nsubsets;
        }
        int s0 = t0 + l * Ntper;
        int s1 = s0 + Ntper;
        if (s0 > t1) s0 = t1;
        if (s1 > t1 || l == Nsubsets - 1) s1 = t1;
        // J^T J and J^T e restricted to the subset; e is the current residual d - f(p)
        gram(pr, k, s0, s1, 1, w.Tsub);
        db_cluster_pass(pr, k, pblk_dev, w.dbuf, nullptr, 1, 0, w.JTe, 2, s0, s1);
        DB_CHECK(cudaMemcpyAsync(hjte, w.JTe, sizeof(double) * n, cudaMemcpyDeviceToHost,
                                 d.stream));
        Tuse = w.Tsub;
      }
      assemble(pr, Tuse, pblk_dev, w.JTJ0);
      DB_CHECK(cudaMemcpyAsync(hH, w.Hst, sizeof(double) * 4 * d.N, cudaMemcpyDeviceToHost,
                               d.stream));
      DB_CHECK(cudaStreamSynchronize(d.stream));
      jacTe_inf = 0.0;
      for (int i = 0; i < n; i++) {
        double a = fabs(hjte[i]);
        if (a > jacTe_inf) jacTe_inf = a;
      }
      const double p_L2 = nrm2sq(hp, n);
      if (jacTe_inf <= eps1) {
        Dp_L2 = 0.0;
        stop = 1;
        break;
      }
      if (kiter == 0 && ositer == 0) {
        // mu0 = tau * max_i (J^T J)_ii ; the diagonal is (h00 x4, h11 x4) per station
        double mx = 0.0;
        for (int s = 0; s < d.N; s++) {
          if (fabs(hH[4 * s]) > fabs(mx)) mx = hH[4 * s];
          if (fabs(hH[4 * s + 1]) > fabs(mx)) mx = hH[4 * s + 1];
        }
        mu = tau * mx;
      }
      // adaptive damping loop (clmfit.c:356-540)
      while (1) {
        int issolved = damped_solve(pr, mu, linsolv, eps1);
        if (linsolv == 2) {
          DB_CHECK(cudaMemcpyAsync(hDp, w.Dp, sizeof(double) * n, cudaMemcpyDeviceToHost,
                                   d.stream));
          DB_CHECK(cudaStreamSynchronize(d.stream));
        }
        if (issolved) {
          for (int i = 0; i < n; i++) hpnew[i] = hp[i] + hDp[i];
          Dp_L2 = nrm2sq(hDp, n);
          if (Dp_L2 <= eps2_sq * p_L2) {
            stop = 2;
            break;
          }
          if (Dp_L2 >= (p_L2 + eps2) / (1e-12 * 1e-12)) {  // CLM_EPSILON, Dirac_common.h:45
            stop = 4;
            break;
          }
          DB_CHECK(cudaMemcpyAsync(w.pnew, hpnew, sizeof(double) * n, cudaMemcpyHostToDevice,
                                   d.stream));
          // trial residual norm and, speculatively, J^T e at the trial point
          db_cluster_pass(pr, k, w.pnew, w.dbuf, nullptr, 1, 0, os ? nullptr : w.JTe_new, 1, t0,
                          t1);
          const double pDp_eL2 = db_read_scalar(pr, 1);
          if (!isfinite(pDp_eL2)) {
            stop = 7;
            break;
          }
          double dL = 0.0;
          for (int i = 0; i < n; i++) dL += hDp[i] * (mu * hDp[i] + hjte[i]);
          const double dF = p_eL2 - pDp_eL2;
          if (dL > 0.0 && dF > 0.0) {
            double tmp = (2.0 * dF / dL - 1.0);
            tmp = 1.0 - tmp * tmp * tmp;
            mu = mu * ((tmp >= 0.3333333334) ? tmp : 0.3333333334);  // CLM_ONE_THIRD
            nu = 2;
            memcpy(hp, hpnew, sizeof(double) * n);
            DB_CHECK(cudaMemcpyAsync(pblk_dev, w.pnew, sizeof(double) * n,
                                     cudaMemcpyDeviceToDevice, d.stream));
            if (!os) {
              DB_CHECK(cudaMemcpyAsync(hjte, w.JTe_new, sizeof(double) * n,
                                       cudaMemcpyDeviceToHost, d.stream));
              double *t = w.JTe; w.JTe = w.JTe_new; w.JTe_new = t;
              DB_CHECK(cudaStreamSynchronize(d.stream));
            }
            p_eL2 = pDp_eL2;
            break;
          }
        }
        mu *= (double)nu;
        nu2 = nu << 1;
        if (nu2 <= nu) {
          stop = 5;
          break;
        }
        nu = nu2;
      }
      if (stop) break;
    }
  }
  if (kiter >= itmax) stop = 3;
  // residual of the chunk with the final Jones: r = d - f(p)   (lmfit.c:980-981)
  db_cluster_pass(pr, k, pblk_dev, w.dbuf, r, 3, 1, nullptr, 1, t0, t1);
  if (info) {
    info[0] = init_p_eL2;
    info[1] = p_eL2;
    info[2] = jacTe_inf;
    info[3] = Dp_L2;
    info[4] = mu;
    info[5] = (double)kiter;
    info[6] = (double)stop;
    info[7] = info[8] = info[9] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// thin C-ABI: normal equations of one (cluster, chunk) against caller-supplied hidden data
// ------------------------------------------------------------------------------------------------
extern "C" double dirac_b200_normal_eq(dirac_b200_problem *pr, int clus, int chunk,
                                       const double *pblk, const double *xd, double *JTJ,
                                       double *JTe) {
  DevProblem &d = pr->d;
  db_lm_init(pr);
  LMWork &w = pr->lm;
  const int n = w.n8;
  int t0, t1;
  chunk_range(d, clus, chunk, &t0, &t1);
  db_upload_vis(pr, xd, w.dbuf);
  DB_CHECK(cudaMemcpyAsync(w.pnew, pblk, sizeof(double) * n, cudaMemcpyHostToDevice, d.stream));
  db_cluster_pass(pr, clus, w.pnew, w.dbuf, nullptr, 1, 0, w.JTe, 1, t0, t1);
  gram(pr, clus, t0, t1, 1, w.Tsub);
  assemble(pr, w.Tsub, w.pnew, w.JTJ0);
  double c = db_read_scalar(pr, 1);
  if (JTe) DB_CHECK(cudaMemcpy(JTe, w.JTe, sizeof(double) * n, cudaMemcpyDeviceToHost));
  if (JTJ)
    DB_CHECK(cudaMemcpy(JTJ, w.JTJ0, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost));
  DB_CHECK(cudaGetLastError());
  return c;
}
