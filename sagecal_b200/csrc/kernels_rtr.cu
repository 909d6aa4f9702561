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
  // the next row's 8 loads are in flight while the current row's ~250 DFMA run (one thread owns a
  // baseline for all its timeslots: without the prefetch a warp has nothing to hide the latency with)
  double2 Cn[4], dn[4];
  unsigned char fn = 1;
  if (t_lo < t_hi) {
    const long long r = (long long)t_lo * a.Nbase + b;
    fn = a.flag[r];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      Cn[c] = ld_stream(a.coh_k + (long long)c * a.R + r);
      dn[c] = ld_stream(a.d + (long long)c * a.R + r);
    }
  }
  for (int t = t_lo; t < t_hi; t++) {
    double2 C[4], dd[4];
    const unsigned char f = fn;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      C[c] = Cn[c];
      dd[c] = dn[c];
    }
    if (t + 1 < t_hi) {
      const long long r = (long long)(t + 1) * a.Nbase + b;
      fn = a.flag[r];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        Cn[c] = ld_stream(a.coh_k + (long long)c * a.R + r);
        dn[c] = ld_stream(a.d + (long long)c * a.R + r);
      }
    }
    if (f) continue;
    rtr_acc_row(A, C, dd, weighted, Gp, Gq, a.nu, a.tensors != 0);
  }
  const size_t sl = blockIdx.y;
  double *sc = a.sc + sl * 3 * (size_t)a.Nbase;
  sc[b] = A.c0;
  sc[(size_t)a.Nbase + b] = A.slw;
  sc[2 * (size_t)a.Nbase + b] = A.cnt;
  if (!a.tensors) return;
  double2 *TD = a.TD + (sl * (size_t)a.Nbase + b) * 32;  // one baseline = 512 contiguous bytes
  double2 T[16], D[16];
  rtr_acc_expand(A, T, D);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    TD[i] = T[i];
    TD[16 + i] = D[i];
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
// k_rtr_eval: CTA = station s; every baseline end of s is spread over 16 lanes (one tensor entry
// [(ab),(mj)] each, rtr_math.cuh: rtr_lane_terms), two ends per warp.  o > s: baseline (s,o), s plays
// p; o < s: baseline (o,s), s plays q.  The lanes' terms are summed over the ends in registers, over
// the half-warps by one shuffle, over the warps in shared memory in a fixed order, and folded (4
// lanes per 2x2 entry) once per station: no atomics, bit-reproducible.  Every baseline is visited from
// both ends.  The dependent chain per thread is ~40 DFMA (12 complex MACs for a Hessian product)
// instead of ~1000 when one thread evaluated a whole end.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ const RtrEvalArgs &rtr_base(const RtrEvalArgs &P) { return P; }
__device__ __forceinline__ const RtrEvalArgs &rtr_base(const RtrEvalInl &P) { return P.a; }
__device__ __forceinline__ const double2 *rtr_x(const RtrEvalArgs &P) {
  return reinterpret_cast<const double2 *>(P.x);
}
__device__ __forceinline__ const double2 *rtr_x(const RtrEvalInl &P) {
  return reinterpret_cast<const double2 *>(P.xin);
}
__device__ __forceinline__ const double2 *rtr_e(const RtrEvalArgs &P) {
  return reinterpret_cast<const double2 *>(P.eta);
}
__device__ __forceinline__ const double2 *rtr_e(const RtrEvalInl &P) {
  return reinterpret_cast<const double2 *>(P.ein);
}

// ARGS = RtrEvalArgs: Jones / tangent vector in device memory; RtrEvalInl: in the parameter block
// (read through the constant bank, no copy in front of the launch)
template <class ARGS>
__global__ void __launch_bounds__(512) k_rtr_eval(const __grid_constant__ ARGS P) {
  const RtrEvalArgs &a = rtr_base(P);
  const int s = blockIdx.x;
  const int lane16 = threadIdx.x & 15, grp = threadIdx.x >> 4, ngrp = blockDim.x >> 4;
  const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int mj = lane16 & 3, ab = lane16 >> 2;
  const int la = ab >> 1, lb = ab & 1;
  const bool hess = a.eta != nullptr;
  const double2 *X = rtr_x(P);
  const double2 *Et = rtr_e(P);
  const bool want_cost = a.cost != nullptr, want_vec = a.out != nullptr;
  double2 tP = make_double2(0, 0), tQ = make_double2(0, 0);
  double cost = 0.0, cnt = 0.0;
  for (int o = grp; o < a.N; o += ngrp) {
    if (o == s) continue;
    const bool sp = s < o;  // s plays p
    const int p = sp ? s : o, q = sp ? o : s;
    const size_t b = (size_t)baseline_index(p, q, a.N);
    if (a.count && lane16 == 0) cnt += a.sc[2 * (size_t)a.Nbase + b];
    if (!want_vec && !(want_cost && sp)) continue;
    // row a of the p station's Jones, row b of the q station's (and of the tangent vector)
    double2 Gpa[2], Gqb[2], Epa[2], Eqb[2];
    Gpa[0] = X[4 * p + 2 * la];
    Gpa[1] = X[4 * p + 2 * la + 1];
    Gqb[0] = X[4 * q + 2 * lb];
    Gqb[1] = X[4 * q + 2 * lb + 1];
    if (hess) {
      Epa[0] = Et[4 * p + 2 * la];
      Epa[1] = Et[4 * p + 2 * la + 1];
      Eqb[0] = Et[4 * q + 2 * lb];
      Eqb[1] = Et[4 * q + 2 * lb + 1];
    }
    double2 Tc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) Tc[k] = a.TD[b * 32 + (4 * k + mj)];
    const double2 Dv = a.TD[b * 32 + (16 + 4 * ab + mj)];
    if (want_cost && sp && lane16 == 0) cost += a.sc[b];
    double2 term = make_double2(0, 0);
    rtr_lane_terms(lane16, sp, Gpa, Gqb, Epa, Eqb, Tc, Dv, hess, want_cost, want_vec, &term, &cost);
    if (sp) tP = cadd(tP, term);
    else tQ = cadd(tQ, term);
  }
  // the two half-warps (two ends) of a warp, then the warps in order
  __shared__ double2 shP[16][16], shQ[16][16];
  __shared__ double shc[16][2];
  tP.x += __shfl_xor_sync(0xffffffffu, tP.x, 16);
  tP.y += __shfl_xor_sync(0xffffffffu, tP.y, 16);
  tQ.x += __shfl_xor_sync(0xffffffffu, tQ.x, 16);
  tQ.y += __shfl_xor_sync(0xffffffffu, tQ.y, 16);
  cost = warp_sum(cost);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) < 16) {
    shP[warp][lane16] = tP;
    shQ[warp][lane16] = tQ;
  }
  if ((threadIdx.x & 31) == 0) {
    shc[warp][0] = cost;
    shc[warp][1] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 16) {  // lane sums over the warps
    double2 p = make_double2(0, 0), q = make_double2(0, 0);
    for (int w = 0; w < nwarp; w++) {
      p = cadd(p, shP[w][threadIdx.x]);
      q = cadd(q, shQ[w][threadIdx.x]);
    }
    shP[0][threadIdx.x] = p;
    shQ[0][threadIdx.x] = q;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    if (a.out) {
      const double2 v = rtr_lane_fold(shP[0], shQ[0], (int)threadIdx.x);
      a.out[8 * (size_t)s + 2 * threadIdx.x] = v.x;
      a.out[8 * (size_t)s + 2 * threadIdx.x + 1] = v.y;
    }
  } else if (threadIdx.x == 4) {
    double c = 0.0, n = 0.0;
    for (int w = 0; w < nwarp; w++) {
      c += shc[w][0];
      n += shc[w][1];
    }
    if (a.cost) a.cost[s] = c;
    if (a.count) a.count[s] = n;
  }
  if (a.flag) {
    // publish: results are fenced to device scope before the CTA checks in; the last CTA in copies
    // all of them to the host mailbox, fences system-wide once and raises the flag
    __shared__ unsigned int last;
    if (threadIdx.x < 5) __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(a.arrive, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (last) {
      __threadfence();
      const int n8 = 8 * a.N;
      // [8N | N | N] device -> host-mapped, only the parts this launch produced
      for (int i = threadIdx.x; i < n8 + 2 * a.N; i += blockDim.x) {
        const bool live = i < n8 ? (a.out != nullptr)
                                 : (i < n8 + a.N ? (a.cost != nullptr) : (a.count != nullptr));
        if (!live) continue;
        const double *src = i < n8 ? a.out + i
                                   : (i < n8 + a.N ? a.cost + (i - n8) : a.count + (i - n8 - a.N));
        a.hmail[i] = __ldcg(src);
      }
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        *a.arrive = 0u;
        *reinterpret_cast<volatile unsigned long long *>(a.flag) = a.epoch;
      }
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
static int rtr_eval_warps(int N) {
  int nwarp = (N + 1) / 2;     // 16 lanes per baseline end, two ends per warp: half of the ends at once
  if (nwarp > 16) nwarp = 16;  // at 62 stations, a loop beyond
  return nwarp < 1 ? 1 : nwarp;
}
void db_launch_rtr_eval(const RtrEvalArgs *a, cudaStream_t st) {
  k_rtr_eval<RtrEvalArgs><<<a->N, 32 * rtr_eval_warps(a->N), 0, st>>>(*a);
}
void db_launch_rtr_eval_inl(const RtrEvalInl *a, cudaStream_t st) {
  k_rtr_eval<RtrEvalInl><<<a->a.N, 32 * rtr_eval_warps(a->a.N), 0, st>>>(*a);
}
void db_launch_rtr_plane_sum(const double *sc, int Nbase, int which, double *dst, cudaStream_t st) {
  k_rtr_plane_sum<<<1, 256, 0, st>>>(sc, Nbase, which, dst);
}
}
