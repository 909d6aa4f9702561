// Robust (Student's-t, IRLS) LM — placeholder until the weighted normal-equation kernel lands.
#include "problem.h"
void db_rlm_chunk(dirac_b200_problem *pr, int k, int ck, double *pblk_dev, double2 *r, int itmax,
                  int linsolv, int os, int randomize, double nulow, double nuhigh,
                  double *robust_nu, double *info) {
  (void)pr; (void)k; (void)ck; (void)pblk_dev; (void)r; (void)itmax; (void)linsolv; (void)os;
  (void)randomize; (void)nulow; (void)nuhigh; (void)robust_nu; (void)info;
  fprintf(stderr, "dirac_b200: robust LM (solver_mode 2/3, last EM iteration) not implemented yet\n");
  exit(1);
}
