// Internal: argument blocks of the RTR-family kernels (kernels_rtr.cu) and their launchers.
#pragma once
#include "internal.cuh"

struct RtrStatsArgs {
  const double2 *coh_k;       // [4][R] coherencies of the cluster
  const double2 *d;           // [4][R] hidden data
  const unsigned char *flag;  // [R]
  const short2 *blpq;         // [Nbase]
  const double *xw;           // 8N Jones the robust weights are evaluated at; null: w = 1
  double nu;
  double2 *TD;                // [nslice][Nbase][32]: per baseline 0-15 T (row-major 4x4), 16-31 D
  double *sc;                 // [nslice][3][Nbase]: c0, sum(log w - w), unflagged rows
  long long R;
  int N, Nbase;
  int t_begin, t_end, tslice;
  int tensors;                // 0: scalars only (weight statistics for the nu update)
};

struct RtrEvalArgs {
  const double2 *TD;          // [Nbase][32]
  const double *sc;           // [3][Nbase]
  const double *x;            // 8N Jones
  const double *eta;          // 8N tangent vector: Hessian-vector product; null: gradient
  double *out;                // [8N] raw station sums (before scaling / projection), may be null
  double *cost;               // [N] per-station partial costs, may be null
  double *count;              // [N] unflagged rows per station, may be null
  int N, Nbase;
  // mailbox: out / cost / count are DEVICE buffers laid out back to back [8N | N | N]; the last CTA
  // to finish copies them to the host-mapped `hmail`, fences system-wide ONCE and publishes `epoch` in
  // *flag (host-mapped too); the host spins on the flag instead of waiting for a device-to-host copy
  // and a stream synchronisation
  unsigned int *arrive;       // device counter, zero between launches
  double *hmail;              // host-mapped [8N + 2N]
  unsigned long long *flag;   // null: no mailbox
  unsigned long long epoch;
};

// up to RTR_INLINE_MAXN stations the Jones and the tangent vector travel in the kernel's parameter
// block (2 x 4 KB; CUDA 12.1+ takes 32 KB of parameters): an evaluation is then ONE launch, no
// host-to-device copy in front of it
#define RTR_INLINE_MAXN 64
struct RtrEvalInl {
  RtrEvalArgs a;   // a.x / a.eta are ignored (a.eta only says: Hessian product or gradient)
  double xin[8 * RTR_INLINE_MAXN];
  double ein[8 * RTR_INLINE_MAXN];
};

extern "C" {
void db_launch_rtr_eval_inl(const RtrEvalInl *a, cudaStream_t st);
void db_launch_rtr_stats(const RtrStatsArgs *a, int nslice, cudaStream_t st);
void db_launch_rtr_reduce(const double *in, double *out, size_t n, int ns, cudaStream_t st);
void db_launch_rtr_eval(const RtrEvalArgs *a, cudaStream_t st);
void db_launch_rtr_plane_sum(const double *sc, int Nbase, int which, double *dst, cudaStream_t st);
}
