/* TEST INFRASTRUCTURE ONLY (oracle).  Compiles the reference's src/lib/Dirac/lmfit.c UNMODIFIED by
 * inclusion (the Makefile passes -I to where it lies under /root/reference; nothing is copied) and
 * exports thin wrappers around its file-static LM callbacks so that the parity tests can call the
 * reference's own predict / Jacobian code directly:
 *   mylm_fit_single_pth   lmfit.c:137   (one cluster, hybrid aware)
 *   mylm_fit_single_pth0  lmfit.c:300   (LM func: one cluster, one chunk)
 *   mylm_jac_single_pth   lmfit.c:484   (LM jacf: dense row-major Jacobian)
 */
#include "lmfit.c"

size_t ref_sizeof_me_data(void) { return sizeof(me_data_t); }

void ref_fill_me_data(void *buf, int clus, int Nbase, int tilesz, int N, baseline_t *barr,
                      clus_source_t *carr, int M, int Mt, double *freq0, int Nt,
                      complex double *coh, int tileoff, double robust_nu) {
  me_data_t *d = (me_data_t *)buf;
  memset(d, 0, sizeof(me_data_t));
  d->clus = clus; d->Nbase = Nbase; d->tilesz = tilesz; d->N = N; d->barr = barr; d->carr = carr;
  d->M = M; d->Mt = Mt; d->freq0 = freq0; d->Nt = Nt; d->coh = coh; d->tileoff = tileoff;
  d->robust_nu = robust_nu;
}
double ref_get_robust_nu(void *buf) { return ((me_data_t *)buf)->robust_nu; }

void ref_mylm_fit_single_pth(double *p, double *x, int m, int n, void *data) {
  mylm_fit_single_pth(p, x, m, n, data);
}
void ref_mylm_fit_single_pth0(double *p, double *x, int m, int n, void *data) {
  mylm_fit_single_pth0(p, x, m, n, data);
}
void ref_mylm_jac_single_pth(double *p, double *jac, int m, int n, void *data) {
  mylm_jac_single_pth(p, jac, m, n, data);
}
