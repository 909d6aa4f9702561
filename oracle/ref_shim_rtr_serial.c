/* TEST INFRASTRUCTURE ONLY (oracle).  Compiles the reference's src/lib/Dirac/rtr_solve_robust.c
 * UNMODIFIED by inclusion, with its worker threads run synchronously inside pthread_create.
 * Why: fns_fupdate_weights adds up its threads' partial sums BEFORE joining them
 * (rtr_solve_robust.c:361-370), so the nu the threaded build returns depends on thread timing
 * (observed here: the top of the nu grid on most runs, the intended value on small chunks).  Run
 * serially, every sum is complete when it is read: this build is the deterministic pin of the robust
 * RTR / NSD solvers (solver_mode 5, 6).  Everything else in _ref/libdirac_ref_serial.so is the same
 * object code as _ref/libdirac_ref.so. */
#include <pthread.h>
static int ser_create(pthread_t *t, const pthread_attr_t *a, void *(*fn)(void *), void *arg) {
  (void)t; (void)a;
  fn(arg);
  return 0;
}
#define pthread_create ser_create
#define pthread_join(t, r) 0
#ifdef SHIM_ADMM
#include "rtr_solve_robust_admm.c" /* same race, rtr_solve_robust_admm.c:381-390 */
#else
#include "rtr_solve_robust.c"
#endif
