/* A plain C host of libdirac_b200.so: compiled against include/dirac_b200.h and LINKED (not dlopened)
 * against the library, the way the reference driver would be.  Without arguments it only calls the
 * host-side helpers (no GPU needed); with "gpu" it also runs sagefit_visibilities on a tiny problem
 * whose coherencies it predicts through precalculate_coherencies. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dirac_b200.h"

int main(int argc, char **argv) {
  const int N = 6, Nbase = N * (N - 1) / 2, tilesz = 4, M = 2;
  const int R = Nbase * tilesz;
  baseline_t *barr = (baseline_t *)calloc(R, sizeof(baseline_t));
  if (generate_baselines(Nbase, tilesz, N, barr, 2) != 0) return 2;
  for (int t = 0; t < tilesz; t++) {
    int b = 0;
    for (int p = 0; p < N - 1; p++)
      for (int q = p + 1; q < N; q++, b++)
        if (barr[t * Nbase + b].sta1 != p || barr[t * Nbase + b].sta2 != q) return 3;
  }
  double *x = (double *)calloc((size_t)8 * R, sizeof(double));
  double *flag = (double *)calloc(R, sizeof(double));
  for (int i = 0; i < 8 * R; i++) x[i] = 1.0 + i;
  flag[3] = 1.0;
  preset_flags_and_data(R, flag, barr, x, 2);
  if (barr[3].flag != 1 || x[8 * 3 + 5] != 0.0 || x[8 * 4] == 0.0) return 4;
  /* every entry point resolves at link time: take their addresses */
  void *syms[] = {(void *)sagefit_visibilities, (void *)sagefit_visibilities_dual_pt_flt,
                  (void *)bfgsfit_visibilities, (void *)bfgsfit_visibilities_gpu,
                  (void *)precalculate_coherencies, (void *)predict_visibilities_multifreq,
                  (void *)calculate_residuals_multifreq, (void *)sagefit_visibilities_admm,
                  (void *)dirac_b200_nccl_init, (void *)dirac_b200_consensus_step};
  for (unsigned i = 0; i < sizeof(syms) / sizeof(syms[0]); i++)
    if (!syms[i]) return 5;
  double B[6];
  double freqs[2] = {140e6, 160e6};
  if (dirac_b200_consensus_basis(B, 3, 2, freqs, 150e6, 0) != 0 || B[0] != 1.0) return 6;
  if (argc > 1 && !strcmp(argv[1], "gpu")) {
    /* two point sources, one per cluster */
    clus_source_t carr[2];
    double ll[2] = {0.01, -0.02}, mm[2] = {0.005, 0.01}, nn[2], sI[2] = {2.0, 1.0}, zero[2] = {0, 0};
    double f0[2] = {150e6, 150e6};
    unsigned char st[2] = {STYPE_POINT, STYPE_POINT};
    int poff[2] = {0, 8 * N};
    void *ex[2] = {0, 0};
    memset(carr, 0, sizeof(carr));
    for (int k = 0; k < M; k++) {
      nn[k] = sqrt(1.0 - ll[k] * ll[k] - mm[k] * mm[k]) - 1.0;
      carr[k].N = 1; carr[k].id = k;
      carr[k].ll = &ll[k]; carr[k].mm = &mm[k]; carr[k].nn = &nn[k];
      carr[k].sI = &sI[k]; carr[k].sQ = &zero[k]; carr[k].sU = &zero[k]; carr[k].sV = &zero[k];
      carr[k].ra = &zero[k]; carr[k].dec = &zero[k]; carr[k].stype = &st[k]; carr[k].ex = &ex[k];
      carr[k].nchunk = 1; carr[k].p = &poff[k];
      carr[k].sI0 = &sI[k]; carr[k].sQ0 = &zero[k]; carr[k].sU0 = &zero[k]; carr[k].sV0 = &zero[k];
      carr[k].f0 = &f0[k]; carr[k].spec_idx = &zero[k]; carr[k].spec_idx1 = &zero[k];
      carr[k].spec_idx2 = &zero[k];
    }
    double *u = (double *)calloc(R, 8), *v = (double *)calloc(R, 8), *w = (double *)calloc(R, 8);
    for (int r = 0; r < R; r++) {
      u[r] = 1e-6 * sin(0.37 * r + 1.0);
      v[r] = 1e-6 * cos(0.11 * r);
      w[r] = 1e-8 * sin(0.05 * r);
      barr[r].flag = 0;
    }
    double *coh = (double *)calloc((size_t)8 * M * R, sizeof(double));
    precalculate_coherencies(u, v, w, coh, N, R, barr, carr, M, 150e6, 180e3, 10.0, 1.0, 0.0, 1e9, 2);
    /* data = model with Jones 1.1 I, start from the identity */
    double *pp = (double *)calloc((size_t)8 * N * M, sizeof(double));
    for (int r = 0; r < R; r++)
      for (int c = 0; c < 8; c++) {
        double s = 0.0;
        for (int k = 0; k < M; k++) s += 1.21 * coh[((size_t)r * M + k) * 8 + c];
        x[8 * r + c] = s;
      }
    for (int k = 0; k < M; k++)
      for (int s = 0; s < N; s++) pp[8 * (k * N + s)] = pp[8 * (k * N + s) + 6] = 1.0;
    double nu, r0, r1;
    int rc = sagefit_visibilities(u, v, w, x, N, Nbase, tilesz, barr, carr, coh, M, M, 150e6, 180e3,
                                  pp, 0.0, 2, 4, 4, 10, 5, 64, 0, 1, 2.0, 30.0, 0, &nu, &r0, &r1);
    printf("sagefit rc=%d res %.3e -> %.3e J00=%.6f\n", rc, r0, r1, pp[0]);
    if (rc != 0 || !(r1 < 1e-3 * r0)) return 7;
  }
  printf("C_CALLER OK\n");
  return 0;
}
