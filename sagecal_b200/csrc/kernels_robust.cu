// Kernels of the robust (Student's-t, IRLS) LM: weighted normal matrix, weight / nu update.
//
// The reference scales every Jacobian row by its sqrt-weight (robustlm.c:2298-2307) and calls dgemm.
// With separate weights on the real and imaginary part of each visibility component c = (i,j),
// a = w_re^2, b = w_im^2, s = (a+b)/2, dl = (a-b)/2, X = C Jq^H, Y = Jp C, the 2x2 sub-blocks are
//   (p,q)[(i,l),(j,l')] = R(Z1) S + R(conj Z2),   Z1 = sum s conj(X_lj) Y_il',  Z2 = sum dl X_lj Y_il'
//   (p,p)[(i,l),(i,l')] = R(P1) + T(P2),          P1 = sum_j s conj(X_lj) X_l'j, P2 = sum_j dl X_lj X_l'j
//   (q,q)[(j,l),(j,l')] = R(conj Q1) + T'(Q2),    Q1 = sum_i s conj(Y_il) Y_il', Q2 = sum_i dl Y_il Y_il'
// with R(z) = [[zr,-zi],[zi,zr]], S = diag(1,-1), T(z) = [[zr,-zi],[-zi,-zr]], T'(z) = [[zr,zi],[zi,-zr]]
// (DESIGN.md "weighted normal equations").  For unit weights this reduces to the Gram-tensor form of
// kernels_lm.cu.  One polarisation product c per sweep of the CTA over its rows keeps the live
// accumulators at 36 doubles per thread; sweeps 2-4 re-read the CTA's rows from L2.
#include "internal.cuh"

#define NSV 10  // station-sum values per (station, sub-block)

__global__ void __launch_bounds__(TILE_THREADS)
k_weighted_jtj(WeightedJtjArgs a) {
  __shared__ double sq[TILE_P][NSV][TILE_Q];
  const TileDesc td = a.tiles[blockIdx.x];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = td.pb * TILE_P + w;
  const int q = td.qb * TILE_Q + lane;
  const bool valid = (q > p) && (q < a.N);
  const int ts = a.t_begin + blockIdx.y * a.tslice;
  const int te = min(ts + a.tslice, a.t_end);
  const int ld = 8 * a.N;
  double2 Jp[4], Jq[4];
#pragma unroll
  for (int c = 0; c < 4; c++) Jp[c] = Jq[c] = make_double2(0.0, 0.0);
  long long b = 0;
  if (valid) {
    load_jones(a.pblk, p, Jp);
    load_jones(a.pblk, q, Jq);
    b = baseline_index(p, q, a.N);
  }
  for (int c = 0; c < 4; c++) {
    const int i = c >> 1, j = c & 1;
    double2 Z1[4], Z2[4], P1o, P2[3], Q1o, Q2[3];
    double P1d[2] = {0.0, 0.0}, Q1d[2] = {0.0, 0.0};
#pragma unroll
    for (int z = 0; z < 4; z++) Z1[z] = Z2[z] = make_double2(0.0, 0.0);
#pragma unroll
    for (int z = 0; z < 3; z++) P2[z] = Q2[z] = make_double2(0.0, 0.0);
    P1o = Q1o = make_double2(0.0, 0.0);
    if (valid) {
      for (int t = ts; t < te; t++) {
        const long long row = (long long)t * a.Nbase + b;
        if (a.flag[row] != 0) continue;
        double2 C[4];
#pragma unroll
        for (int z = 0; z < 4; z++) C[z] = ld_stream(a.coh_k + (long long)z * a.R + row);
        const double2 wv = ld_stream(a.wt + (long long)c * a.R + row);
        const double wa = wv.x * wv.x, wb = wv.y * wv.y;
        const double s = 0.5 * (wa + wb), dl = 0.5 * (wa - wb);
        // X_l = X[l][j] = sum_a C[l][a] conj(Jq[j][a]) ;  Y_l' = Y[i][l'] = sum_b Jp[i][b] C[b][l']
        double2 X[2], Y[2];
#pragma unroll
        for (int l = 0; l < 2; l++) {
          X[l] = cadd(cmulc(C[2 * l], Jq[2 * j]), cmulc(C[2 * l + 1], Jq[2 * j + 1]));
          Y[l] = cadd(cmul(Jp[2 * i], C[l]), cmul(Jp[2 * i + 1], C[2 + l]));
        }
#pragma unroll
        for (int l = 0; l < 2; l++)
#pragma unroll
          for (int lp = 0; lp < 2; lp++) {
            const double2 u1 = cmulcl(X[l], Y[lp]);  // conj(X_l) Y_l'
            const double2 u2 = cmul(X[l], Y[lp]);
            Z1[2 * l + lp].x = fma(s, u1.x, Z1[2 * l + lp].x);
            Z1[2 * l + lp].y = fma(s, u1.y, Z1[2 * l + lp].y);
            Z2[2 * l + lp].x = fma(dl, u2.x, Z2[2 * l + lp].x);
            Z2[2 * l + lp].y = fma(dl, u2.y, Z2[2 * l + lp].y);
          }
#pragma unroll
        for (int l = 0; l < 2; l++) {
          P1d[l] = fma(s, X[l].x * X[l].x + X[l].y * X[l].y, P1d[l]);
          Q1d[l] = fma(s, Y[l].x * Y[l].x + Y[l].y * Y[l].y, Q1d[l]);
        }
        {
          const double2 u = cmulcl(X[0], X[1]), v = cmulcl(Y[0], Y[1]);
          P1o.x = fma(s, u.x, P1o.x); P1o.y = fma(s, u.y, P1o.y);
          Q1o.x = fma(s, v.x, Q1o.x); Q1o.y = fma(s, v.y, Q1o.y);
          const double2 x00 = cmul(X[0], X[0]), x01 = cmul(X[0], X[1]), x11 = cmul(X[1], X[1]);
          const double2 y00 = cmul(Y[0], Y[0]), y01 = cmul(Y[0], Y[1]), y11 = cmul(Y[1], Y[1]);
          P2[0].x = fma(dl, x00.x, P2[0].x); P2[0].y = fma(dl, x00.y, P2[0].y);
          P2[1].x = fma(dl, x01.x, P2[1].x); P2[1].y = fma(dl, x01.y, P2[1].y);
          P2[2].x = fma(dl, x11.x, P2[2].x); P2[2].y = fma(dl, x11.y, P2[2].y);
          Q2[0].x = fma(dl, y00.x, Q2[0].x); Q2[0].y = fma(dl, y00.y, Q2[0].y);
          Q2[1].x = fma(dl, y01.x, Q2[1].x); Q2[1].y = fma(dl, y01.y, Q2[1].y);
          Q2[2].x = fma(dl, y11.x, Q2[2].x); Q2[2].y = fma(dl, y11.y, Q2[2].y);
        }
      }
      // off-diagonal blocks (p,q) and (q,p): this thread is the only writer within its time slice
#pragma unroll
      for (int l = 0; l < 2; l++)
#pragma unroll
        for (int lp = 0; lp < 2; lp++) {
          const double2 z1 = Z1[2 * l + lp], z2 = Z2[2 * l + lp];
          const double v00 = z1.x + z2.x, v01 = z1.y + z2.y, v10 = z1.y - z2.y, v11 = -z1.x + z2.x;
          const long long r0 = 8 * p + 2 * (2 * i + l), c0 = 8 * q + 2 * (2 * j + lp);
          atomicAdd(a.JTJ + r0 * ld + c0, v00);
          atomicAdd(a.JTJ + r0 * ld + c0 + 1, v01);
          atomicAdd(a.JTJ + (r0 + 1) * ld + c0, v10);
          atomicAdd(a.JTJ + (r0 + 1) * ld + c0 + 1, v11);
          atomicAdd(a.JTJ + c0 * ld + r0, v00);
          atomicAdd(a.JTJ + (c0 + 1) * ld + r0, v01);
          atomicAdd(a.JTJ + c0 * ld + r0 + 1, v10);
          atomicAdd(a.JTJ + (c0 + 1) * ld + r0 + 1, v11);
        }
    }
    // station sums: p-role over the lanes of the warp, q-role over the warps of the CTA
    double pv[NSV] = {P1d[0], P1d[1], P1o.x, P1o.y, P2[0].x, P2[0].y, P2[1].x, P2[1].y, P2[2].x,
                      P2[2].y};
    double qv[NSV] = {Q1d[0], Q1d[1], Q1o.x, Q1o.y, Q2[0].x, Q2[0].y, Q2[1].x, Q2[1].y, Q2[2].x,
                      Q2[2].y};
#pragma unroll
    for (int z = 0; z < NSV; z++) pv[z] = warp_sum(pv[z]);
    if (p < a.N - 1 && lane < NSV) {
      double v = pv[0];
#pragma unroll
      for (int z = 1; z < NSV; z++) v = (lane == z) ? pv[z] : v;
      atomicAdd(a.HP + ((long long)p * 2 + i) * NSV + lane, v);
    }
#pragma unroll
    for (int z = 0; z < NSV; z++) sq[w][z][lane] = qv[z];
    __syncthreads();
    for (int z = w; z < NSV; z += TILE_P) {
      double s = 0.0;
#pragma unroll
      for (int ww = 0; ww < TILE_P; ww++) s += sq[ww][z][lane];
      if (q < a.N && s != 0.0) atomicAdd(a.HQ + ((long long)q * 2 + j) * NSV + z, s);
    }
    __syncthreads();
  }
}

// diagonal 8x8 blocks from the station sums; one thread per station
__global__ void k_weighted_diag(const double *__restrict__ HP, const double *__restrict__ HQ,
                                double *__restrict__ JTJ, int N) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  const int ld = 8 * N;
  for (int i = 0; i < 2; i++) {
    const double *P = HP + ((long long)s * 2 + i) * NSV;
    const double *Q = HQ + ((long long)s * 2 + i) * NSV;
    double blk[4][4];
    for (int l = 0; l < 2; l++)
      for (int lp = 0; lp < 2; lp++) {
        // P1[l][l'], Q1[l][l'] (Hermitian), P2, Q2 (symmetric)
        double p1r, p1i, q1r, q1i;
        if (l == lp) {
          p1r = P[l]; p1i = 0.0; q1r = Q[l]; q1i = 0.0;
        } else if (l < lp) {
          p1r = P[2]; p1i = P[3]; q1r = Q[2]; q1i = Q[3];
        } else {
          p1r = P[2]; p1i = -P[3]; q1r = Q[2]; q1i = -Q[3];
        }
        const int k2 = l + lp;  // (0,0)->0 (0,1),(1,0)->1 (1,1)->2
        const double p2r = P[4 + 2 * k2], p2i = P[5 + 2 * k2];
        const double q2r = Q[4 + 2 * k2], q2i = Q[5 + 2 * k2];
        // R(P1) + T(P2) + R(conj Q1) + T'(Q2)
        blk[2 * l][2 * lp] = p1r + p2r + q1r + q2r;
        blk[2 * l][2 * lp + 1] = -p1i - p2i + q1i + q2i;
        blk[2 * l + 1][2 * lp] = p1i - p2i - q1i + q2i;
        blk[2 * l + 1][2 * lp + 1] = p1r - p2r + q1r - q2r;
      }
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 8; c++) {
        double v = 0.0;
        if ((c >> 2) == i) v = blk[r][c & 3];
        JTJ[(long long)(8 * s + 4 * i + r) * ld + 8 * s + c] = v;
      }
  }
}

// ---- elementwise over the rows [r0, r1) of a planar [4][R] vector -----------------------------------
struct RowRange {
  long long R, r0, r1;
};
__device__ __forceinline__ long long rr_index(const RowRange &g, long long i) {
  const long long nr = g.r1 - g.r0;
  return (i / nr) * g.R + g.r0 + (i % nr);
}

// deterministic sum of |v| over the range -> *out
__global__ void __launch_bounds__(256)
k_sum_abs(const double2 *__restrict__ v, RowRange g, double *partials, double *out,
          unsigned int *counter);

// w <- sqrt((nu+1)/(nu+e^2)); accumulates sum |w - log w| (w before the sqrt)  (updatenu.c:60-78)
__global__ void __launch_bounds__(256)
k_update_weights(const double2 *__restrict__ e, double2 *__restrict__ wt, RowRange g, double nu0,
                 double *partials, double *out, unsigned int *counter);

// the definitions need grid_reduce_sum, which lives in kernels_stream.cu as a static inline; keep a
// private copy here (one per translation unit)
__device__ __forceinline__ void grid_reduce_sum_r(double v, double *partials, double *out,
                                                  unsigned int *counter) {
  __shared__ double wsum[32];
  __shared__ bool is_last;
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) wsum[w] = v;
  __syncthreads();
  const unsigned int nblocks = gridDim.x;
  if (threadIdx.x == 0) {
    double s = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; i++) s += wsum[i];
    partials[blockIdx.x] = s;
    __threadfence();
    is_last = (atomicAdd(counter, 1u) == nblocks - 1);
  }
  __syncthreads();
  if (is_last) {
    double s = 0.0;
    for (unsigned int i = threadIdx.x; i < nblocks; i += blockDim.x)
      s += ((volatile double *)partials)[i];
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) wsum[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      const int nw = (blockDim.x + 31) >> 5;
      for (int i = 0; i < nw; i++) tot += wsum[i];
      *out = tot;
      *counter = 0;
    }
  }
}

__global__ void __launch_bounds__(256)
k_sum_abs(const double2 *__restrict__ v, RowRange g, double *partials, double *out,
          unsigned int *counter) {
  const long long n = 4 * (g.r1 - g.r0);
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double2 x = v[rr_index(g, i)];
    s += fabs(x.x) + fabs(x.y);
  }
  grid_reduce_sum_r(s, partials, out, counter);
}

__global__ void __launch_bounds__(256)
k_update_weights(const double2 *__restrict__ e, double2 *__restrict__ wt, RowRange g, double nu0,
                 double *partials, double *out, unsigned int *counter) {
  const long long n = 4 * (g.r1 - g.r0);
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long ix = rr_index(g, i);
    const double2 ev = e[ix];
    const double wx = (nu0 + 1.0) / (nu0 + ev.x * ev.x);
    const double wy = (nu0 + 1.0) / (nu0 + ev.y * ev.y);
    s += fabs(wx - log(wx)) + fabs(wy - log(wy));
    wt[ix] = make_double2(sqrt(wx), sqrt(wy));
  }
  grid_reduce_sum_r(s, partials, out, counter);
}

__global__ void __launch_bounds__(256)
k_scale_vis(double2 *__restrict__ v, RowRange g, double alpha, int set_const) {
  const long long n = 4 * (g.r1 - g.r0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long ix = rr_index(g, i);
    if (set_const) {
      v[ix] = make_double2(alpha, alpha);
    } else {
      double2 x = v[ix];
      v[ix] = make_double2(alpha * x.x, alpha * x.y);
    }
  }
}

// Misaligned ordered subset (clmfit.c:1313-1413): row i = 8*rho + comp of the subset's Jacobian is paired
// with the chunk's data index kl + i.  eps / wout over the subset's rows rho (planar, absolute rows
// row_sub0 + rho): the residual and the sqrt-weight found at that index, zero beyond the cut nJ.
__global__ void __launch_bounds__(256)
k_os_shift(const double2 *__restrict__ e, const double2 *__restrict__ wt, double2 *__restrict__ eps,
           double2 *__restrict__ wout, long long R, long long row_sub0, long long nrow_sub,
           long long row_chunk0, long long kl, long long nJ) {
  const long long tot = 4 * nrow_sub;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < tot;
       t += (long long)gridDim.x * blockDim.x) {
    const long long rho = t >> 2;
    const int c4 = (int)(t & 3);
    double ev[2], wv[2];
#pragma unroll
    for (int cc = 0; cc < 2; cc++) {
      const long long i = 8 * rho + 2 * c4 + cc;
      if (i < nJ) {
        const long long g = kl + i;
        const long long srow = row_chunk0 + (g >> 3);
        const int sc = (int)(g & 7);
        const double2 se = e[(long long)(sc >> 1) * R + srow];
        ev[cc] = (sc & 1) ? se.y : se.x;
        if (wt) {
          const double2 sw = wt[(long long)(sc >> 1) * R + srow];
          wv[cc] = (sc & 1) ? sw.y : sw.x;
        } else {
          wv[cc] = 1.0;
        }
      } else {
        ev[cc] = 0.0;
        wv[cc] = 0.0;
      }
    }
    eps[(long long)c4 * R + row_sub0 + rho] = make_double2(ev[0], ev[1]);
    wout[(long long)c4 * R + row_sub0 + rho] = make_double2(wv[0], wv[1]);
  }
}

extern "C" {
void db_launch_os_shift(const double2 *e, const double2 *wt, double2 *eps, double2 *wout, long long R,
                        long long row_sub0, long long nrow_sub, long long row_chunk0, long long kl,
                        long long nJ, cudaStream_t st) {
  long long tot = 4 * nrow_sub;
  int grid = (int)((tot + 255) / 256 < 1184 ? (tot + 255) / 256 : 1184);
  k_os_shift<<<grid, 256, 0, st>>>(e, wt, eps, wout, R, row_sub0, nrow_sub, row_chunk0, kl, nJ);
}
void db_launch_weighted_jtj(const WeightedJtjArgs *a, int ntile, cudaStream_t st) {
  const int nt = a->t_end - a->t_begin;
  dim3 grid(ntile, (nt + a->tslice - 1) / a->tslice);
  k_weighted_jtj<<<grid, TILE_THREADS, 0, st>>>(*a);
  k_weighted_diag<<<(a->N + 63) / 64, 64, 0, st>>>(a->HP, a->HQ, a->JTJ, a->N);
}
#define ROBUST_GRID 296
void db_launch_sum_abs(const double2 *v, long long R, long long r0, long long r1, double *partials,
                       double *out, unsigned int *counter, cudaStream_t st) {
  RowRange g = {R, r0, r1};
  k_sum_abs<<<ROBUST_GRID, 256, 0, st>>>(v, g, partials, out, counter);
}
void db_launch_update_weights(const double2 *e, double2 *wt, long long R, long long r0,
                              long long r1, double nu0, double *partials, double *out,
                              unsigned int *counter, cudaStream_t st) {
  RowRange g = {R, r0, r1};
  k_update_weights<<<ROBUST_GRID, 256, 0, st>>>(e, wt, g, nu0, partials, out, counter);
}
void db_launch_scale_vis(double2 *v, long long R, long long r0, long long r1, double alpha,
                         int set_const, cudaStream_t st) {
  RowRange g = {R, r0, r1};
  k_scale_vis<<<ROBUST_GRID, 256, 0, st>>>(v, g, alpha, set_const);
}
}
