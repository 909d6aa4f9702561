// Kernels of the Riemannian trust-region / steepest-descent / Nesterov solvers (solver_mode 4-6).
//
// The reference evaluates cost, gradient and Hessian-vector product of one (cluster, chunk) by a
// pass over all its rows, once per inner iteration (rtr_solve.c:188-315, 453-636, 643-870;
// rtr_solve_robust.c:72-205, 519-716, 722-972).  All three are polynomials in the Jones matrices
// whose coefficients depend on the rows only through per-baseline sums, because the Jones matrices
// are constant over the chunk and the robust weights are one scalar per row, fixed while the
// trust-region loop runs:
//     T[(ij),(kl)] = sum_t w C_ij conj(C_kl)        (4x4 Hermitian)
//     D[(ab),(ij)] = sum_t w d_ab conj(C_ij)        (4x4 complex)
//     c0           = sum_t w |d|^2
// (d hidden data, C coherency, unflagged rows).  With M(A1,A2)[(ab),(mj)] = sum_t w (A1 C A2^H)_ab
// conj(C_mj) = sum_{i,j'} A1_ai conj(A2_bj') T[(ij'),(mj)] and Wres = D - M(Gp,Gq):
//     cost            = c0 - Re sum conj(Gp_ai conj(Gq_bj)) (D + Wres)[(ab),(ij)]
//     grad_p[a,m]     = sum_bj Gq_bj Wres[(ab),(mj)]                  (res Gq C^H)
//     grad_q[b,m]     = sum_ai Gp_ai conj(Wres[(ab),(im)])            (res^H Gp C)
//     hess_p[a,m]     = sum_bj (Eq_bj Wres - Gq_bj W1)[(ab),(mj)],  W1 = M(Gp,Eq) + M(Ep,Gq)
//     hess_q[b,m]     = sum_ai (Ep_ai conj(Wres) - Gp_ai conj(W1))[(ab),(im)]
// So ONE streaming pass (k_rtr_stats, 129 B per row) replaces the ~100 row passes of a visit, and
// every evaluation the solver asks for afterwards costs O(Nbase) (k_rtr_eval).
#include "internal.cuh"
#include "rtr.h"
#include "rtr_math.cuh"

// ------------------------------------------------------------------------------------------------
// k_rtr_stats: thread = baseline (consecutive lanes = consecutive rows of a timeslot), blockIdx.y =
// time slice.  Weights: w = 1, or Student's-t row weights at the Jones xw (rtr_math.cuh).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_rtr_stats(RtrStatsArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.Nbase) return;
  const int t_lo = a.t_begin + blockIdx.y * a.tslice;
  int t_hi = t_lo + a.tslice;
  if (t_hi > a.t_end) t_hi = a.t_end;
  const short2 pq = a.blpq[b];
  double2 Gp[4], Gq[4];
  const bool weighted = a.xw != nullptr;
  if (weighted) {
    load_jones(a.xw, pq.x, Gp);
    load_jones(a.xw, pq.y, Gq);
  }
  RtrAcc A;
  rtr_acc_zero(A);
  for (int t = t_lo; t < t_hi; t++) {
    const long long r = (long long)t * a.Nbase + b;
    if (a.flag[r]) continue;
    double2 C[4], dd[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      C[c] = ld_stream(a.coh_k + (long long)c * a.R + r);
      dd[c] = ld_stream(a.d + (long long)c * a.R + r);
    }
    rtr_acc_row(A, C, dd, weighted, Gp, Gq, a.nu, a.tensors != 0);
  }
  const size_t sl = blockIdx.y;
  double *sc = a.sc + sl * 3 * (size_t)a.Nbase;
  sc[b] = A.c0;
  sc[(size_t)a.Nbase + b] = A.slw;
  sc[2 * (size_t)a.Nbase + b] = A.cnt;
  if (!a.tensors) return;
  double2 *TD = a.TD + sl * 32 * (size_t)a.Nbase;
  double2 T[16], D[16];
  rtr_acc_expand(A, T, D);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    TD[(size_t)i * a.Nbase + b] = T[i];
    TD[(size_t)(16 + i) * a.Nbase + b] = D[i];
  }
}

// out[i] = sum_s in[s][i]  (deterministic sum of the time slices)
__global__ void k_rtr_reduce(const double *in, double *out, size_t n, int ns) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < ns; k++) s += in[(size_t)k * n + i];
  out[i] = s;
}

// ------------------------------------------------------------------------------------------------
// k_rtr_eval: CTA = station s, threads = the other stations o (one baseline per thread up to 257
// stations, a short loop beyond).  o > s: baseline (s,o), s plays p; o < s: baseline (o,s), s plays q.
// Every baseline is visited from both ends: no atomics, and the sums are bit-reproducible.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rtr_eval(RtrEvalArgs a) {
  const int s = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  double2 Gs[4], Es[4];
  load_jones(a.x, s, Gs);
  const bool hess = a.eta != nullptr;
  if (hess) load_jones(a.eta, s, Es);
  double2 acc[4];
#pragma unroll
  for (int i = 0; i < 4; i++) acc[i] = make_double2(0, 0);
  double cost = 0.0, cnt = 0.0;
  for (int o = threadIdx.x; o < a.N; o += blockDim.x) {
    if (o == s) continue;
    const bool sp = s < o;  // s plays p
    const int p = sp ? s : o, q = sp ? o : s;
    const size_t b = (size_t)baseline_index(p, q, a.N);
    if (a.count) cnt += a.sc[2 * (size_t)a.Nbase + b];
    if (!a.out && !(a.cost && sp)) continue;
    double2 Go[4], Eo[4];
    load_jones(a.x, o, Go);
    if (hess) load_jones(a.eta, o, Eo);
    double2 T[16], W[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      T[i] = a.TD[(size_t)i * a.Nbase + b];
      W[i] = a.TD[(size_t)(16 + i) * a.Nbase + b];
    }
    rtr_eval_baseline(sp, Gs, Go, Es, Eo, T, W, a.sc[b], hess, a.cost != nullptr,
                      a.out != nullptr, acc, &cost);
  }
  // 10 sums per station: warp butterflies, then the warps' partials in order through shared memory
  __shared__ double part[8][10];
  double v[10] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y, acc[2].x, acc[2].y, acc[3].x, acc[3].y,
                  cost, cnt};
#pragma unroll
  for (int i = 0; i < 10; i++) {
    v[i] = warp_sum(v[i]);
    if (lane == 0) part[warp][i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double t = 0.0;
    for (int w = 0; w < nwarp; w++) t += part[w][threadIdx.x];
    const int i = threadIdx.x;
    if (i < 8) {
      if (a.out) a.out[8 * (size_t)s + i] = t;
    } else if (i == 8) {
      if (a.cost) a.cost[s] = t;
    } else if (a.count) {
      a.count[s] = t;
    }
    if (a.flag) __threadfence_system();
  }
  if (a.flag) {
    // publish: every CTA's results are fenced out to the host before it checks in; the last one in
    // raises the flag
    __shared__ unsigned int last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      last = (atomicAdd(a.arrive, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
      *a.arrive = 0u;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long *>(a.flag) = a.epoch;
    }
  }
}

// sum of plane `which` (0: c0, 1: sum(log w - w), 2: unflagged rows) over the baselines -> *dst
__global__ void __launch_bounds__(256) k_rtr_plane_sum(const double *sc, int Nbase, int which,
                                                        double *dst) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < Nbase; i += 256) s += sc[(size_t)which * Nbase + i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *dst = sh[0];
}

extern "C" {
void db_launch_rtr_stats(const RtrStatsArgs *a, int nslice, cudaStream_t st) {
  dim3 grid((a->Nbase + 127) / 128, nslice);
  k_rtr_stats<<<grid, 128, 0, st>>>(*a);
}
void db_launch_rtr_reduce(const double *in, double *out, size_t n, int ns, cudaStream_t st) {
  k_rtr_reduce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n, ns);
}
void db_launch_rtr_eval(const RtrEvalArgs *a, cudaStream_t st) {
  int nwarp = (a->N + 31) / 32;  // one baseline per thread where a CTA can hold them
  if (nwarp > 8) nwarp = 8;
  k_rtr_eval<<<a->N, 32 * nwarp, 0, st>>>(*a);
}
void db_launch_rtr_plane_sum(const double *sc, int Nbase, int which, double *dst, cudaStream_t st) {
  k_rtr_plane_sum<<<1, 256, 0, st>>>(sc, Nbase, which, dst);
}
}
