/* TEST INFRASTRUCTURE ONLY (oracle).  Compiles the reference's src/lib/Dirac/robust_lbfgs.c
 * UNMODIFIED by inclusion and exports its file-static LBFGS cost / gradient callbacks:
 *   cost_func         robust_lbfgs.c:674    grad_func         robust_lbfgs.c:697
 *   robust_cost_func  robust_lbfgs.c:707    robust_grad_func  robust_lbfgs.c:729
 */
#include "robust_lbfgs.c"

double ref_cost_func(double *p, int m, double *x, int n, void *medata) {
  wrapper_me_data_t w; w.adata = (me_data_t *)medata; w.x = x; w.n = n; w.func = 0;
  return cost_func(p, m, &w);
}
void ref_grad_func(double *p, double *g, int m, double *x, int n, void *medata) {
  wrapper_me_data_t w; w.adata = (me_data_t *)medata; w.x = x; w.n = n; w.func = 0;
  grad_func(p, g, m, &w);
}
double ref_robust_cost_func(double *p, int m, double *x, int n, void *medata) {
  wrapper_me_data_t w; w.adata = (me_data_t *)medata; w.x = x; w.n = n; w.func = 0;
  return robust_cost_func(p, m, &w);
}
void ref_robust_grad_func(double *p, double *g, int m, double *x, int n, void *medata) {
  wrapper_me_data_t w; w.adata = (me_data_t *)medata; w.x = x; w.n = n; w.func = 0;
  robust_grad_func(p, g, m, &w);
}
