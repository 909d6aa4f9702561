/* TEST ONLY.  The linking recipe of INTEGRATION.md section 2, checked without a GPU: a host linked
 * `-ldirac_b200` BEFORE the reference's own library resolves the hot-path entry points to
 * libdirac_b200.so and everything else (BLAS wrappers, sky-model helpers) to the reference library.
 * Prints "<symbol> <library file>" per line; no compute call is made. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "dirac_b200.h"

/* reference-only symbols (src/lib/Dirac/Dirac.h: my_dscal, my_dnrm2; update_w_and_nu updatenu.c:137) */
extern void my_dscal(int N, double a, double *x);
extern double my_dnrm2(int N, double *x);
extern int update_w_and_nu(int, double *, double *, double *, int, int, double, double, int, double *);

static int where(const char *name, void *fn) {
  Dl_info info;
  if (!dladdr(fn, &info) || !info.dli_fname) {
    printf("%s ?\n", name);
    return 1;
  }
  const char *base = strrchr(info.dli_fname, '/');
  printf("%s %s\n", name, base ? base + 1 : info.dli_fname);
  return 0;
}

#define W(f) bad |= where(#f, (void *)f)
int main(void) {
  int bad = 0;
  W(sagefit_visibilities);
  W(sagefit_visibilities_dual_pt_flt);
  W(bfgsfit_visibilities);
  W(bfgsfit_visibilities_gpu);
  W(precalculate_coherencies);
  W(precalculate_coherencies_withbeam_gpu);
  W(predict_visibilities_multifreq);
  W(calculate_residuals_multifreq);
  W(sagefit_visibilities_admm);
  W(bfgsfit_minibatch_visibilities);
  W(lbfgs_persist_init);
  W(generate_baselines);
  W(preset_flags_and_data);
  W(whiten_data);
  W(my_dscal);
  W(my_dnrm2);
  W(update_w_and_nu);
  return bad;
}
