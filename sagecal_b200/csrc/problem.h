// Internal: the opaque dirac_b200_problem and the device-pointer primitives shared by the solvers.
#pragma once
#include <cusolverDn.h>
#include <cublas_v2.h>

#include "internal.cuh"

struct LMWork {
  bool ready;
  int n8;                 // 8N
  double *T;              // [Mt][Nbase][16] Gram tensors per (cluster, chunk), built on first use
  unsigned char *T_valid; // host, [Mt]
  double *Tsub;           // [Nbase][16] scratch (OS subsets, tests)
  double *JTJ0, *JTJ;     // [8N][8N]
  double *JTe, *JTe_new;  // [8N]
  double *Hst;            // [N][4]
  double *Dp;             // [8N]
  double *pnew;           // [8N]
  double *h_vec;          // pinned host scratch, 4*8N + 4N + 16
  int *devinfo;
  double *cswork;
  int lwork;
  bool own_chol;  // damped solves by the cluster Cholesky kernel (else cuSOLVER)
  double *bt_ws;  // large systems: workspace of the blocked triangular solves (kernels_bigtri.cu)
  unsigned bt_epoch;
  double *jte_part;       // per-CTA station sums of the linear-mapped gradient pass
  bool step_armed, step_fused;  // trial point formed by the solver kernel's epilogue
  double *jtj_spec;       // J^T J assembled speculatively at the trial point (nullptr: none)
  cudaEvent_t ev_mail;    // marks the trial results' copy; the host waits on it, not on the stream
  double *tau;            // QR
  double *svdS, *svdU, *svdVT;
  cusolverDnHandle_t cs;
  cublasHandle_t cb;
  double2 *dbuf;          // [4][R] hidden data of the cluster being solved
  // robust LM (allocated on first use)
  double2 *wbuf;          // [4][R] sqrt-weights
  double2 *ebuf;          // [4][R] unweighted residual for the weight update
  double2 *os_eps, *os_w; // [4][R] misaligned ordered subsets: shifted residual / sqrt-weights + cut
  double *HP, *HQ;        // [N][2][10] station sums of the weighted normal matrix
  double *plast;          // [8N] device copy of the last evaluated trial point
  double *pold;           // [8N] Jones at the start of the visit (sharded closing pass)
  // normal matrices of ALL clusters of a sweep, assembled and factorised in one batch before the
  // sweep (each cluster's first LM solve then only needs the triangular solves)
  double *JB, *LB;        // [M][8N][8N] J^T J and its damped Cholesky factor
  size_t lb_stride;       // doubles between two factors in LB
  int lb_ld, binfo_step;  // leading dimension of a factor; ints per matrix in the status array
  double *HB;             // [M][N][4]
  double *mu_dev;         // [M] mu0 of each cluster
  double *h_mu;           // pinned
  int *binfo_dev, *h_binfo;
  double **LBptr_dev;     // [M] pointers into LB
  int *blist_dev, *btix_dev, *bpoff_dev;
  int *pref_slot;         // host [M]: slot of cluster k in the current batch, -1 if not prefactored
  const double *jtj0_cur; // matrix the damping loop of the current iteration starts from
};

struct dirac_b200_problem {
  DevProblem d;
  double *partials;
  int npartials;
  double2 *res;           // [4][R] residual of the full model
  double2 *vis_stage;     // [R][4] API-layout staging
  double *g;              // [8*N*Mt]
  LMWork lm;
  int own_stream;
  // cluster sharding over ranks: sum of a device buffer of doubles over all ranks (NCCL all-reduce,
  // supplied by the host); world == 1: never called
  int rank, world;
  void (*allreduce)(void *dev, long long count, void *stream, void *user);
  void *comm_user;
  int m_global;           // clusters over all ranks
  int k_global0;          // global index of local cluster 0
  double beta;            // hidden-data weight of the sharded SAGE sweep (1/world by default)
  // consensus (ADMM) terms of the running solve (dirac_b200_sagefit_admm), null otherwise
  const double *aug_y_host, *aug_bz_host, *aug_rho;  // host: [npar], [npar], [M]
  double *aug_dev;        // device: Y | BZ
  double2 *pm;            // [4][R] partial model / residual at the start of a sharded sweep
  double *xb;             // [8R + npar + m_global] sweep exchange message (sharded)
  double *pp_start;       // [npar] Jones at the start of a sharded sweep
  // LBFGS line model (allocated on first use)
  double2 *E0, *E1, *E2;  // [4][R] each
  struct RtrWork *rtr;    // RTR / NSD solvers (rtr.cu), allocated on first use
};

// host waits on the device, timed (dirac_b200_host_stats): where the host-driven solver idles
void db_stream_sync(cudaStream_t st);
void db_event_sync(cudaEvent_t ev);
void db_flag_wait(const volatile unsigned long long *flag, unsigned long long epoch,
                  cudaStream_t st);
void *db_malloc(size_t bytes);
void db_free(void *p);
void db_count_launch(int n);
void db_prof_begin(int kind, double bytes, cudaStream_t st);
void db_prof_end(cudaStream_t st);
cudaStream_t db_new_stream(int *owned);
void db_upload_vis(dirac_b200_problem *pr, const double *h, double2 *dst);
void db_download_vis(dirac_b200_problem *pr, const double2 *src, double *h);
void db_predict_dev(dirac_b200_problem *pr, const double *pp_dev, double2 *out, int out_mode,
                    int cost_mode, double nu, int slot);
double db_read_scalar(dirac_b200_problem *pr, int slot);
void db_grad_dev(dirac_b200_problem *pr, const double *pp_dev, double *g_dev, int robust,
                 double nu);
int db_use_tma();
void db_lm_init(dirac_b200_problem *pr);
void db_prefactor_sweep(dirac_b200_problem *pr, double tau);
void db_allreduce(dirac_b200_problem *pr, void *dev, long long count);
// sum over the ranks of the process-wide communicator regardless of cluster sharding (consensus over
// subbands: every rank holds its own, unsharded problem); no-op without a communicator
void db_allreduce_world(dirac_b200_problem *pr, void *dev, long long count);
void db_lm_set_aug(const double *y_dev, const double *bz_dev, const double *y_host,
                   const double *bz_host, double rho);
int db_overlap_available(const dirac_b200_problem *pr);
cudaStream_t db_comm_stream();
void db_allreduce_segments(dirac_b200_problem *pr, double **ptr, const long long *count, int nseg,
                           cudaStream_t st);
extern "C" {
void db_launch_residual_cost(const double2 *x, const double2 *pm, double2 *out, long long n4,
                             int out_mode, int cost_mode, double inv_nu, double *partials,
                             double *cost, unsigned int *counter, cudaStream_t st);
void db_launch_axpby(const double2 *x, double2 *y, long long n4, double a, double b,
                     cudaStream_t st);
void db_launch_cluster_rowmap(const double2 *coh_k, const double2 *in, const double2 *in2,
                              double2 *out, const unsigned char *flag, const double *pp,
                              const int *chunk_poff, int nchunk, const short2 *blpq, long long R,
                              int Nbase, int sign, double beta, cudaStream_t st);
}

void db_lm_free(dirac_b200_problem *pr);
void db_rtr_free(dirac_b200_problem *pr);
