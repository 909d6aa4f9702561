// Normal-equation assembly and small dense helpers of the per-cluster LM solver.
//
// The reference forms the dense n x 8N Jacobian (jacobian_threadfn, lmfit.c:392-474) and calls
// dgemm for J^T J (clmfit.c:307).  J has 16 non-zeros per row-octet and, for fixed Jones, J^T J
// depends on the data only through the per-baseline Gram tensor Th = sum_t conj(c) c^T of the
// coherencies (k_coh_gram).  With X = C Jq^H, Y = Jp C (V = Jp X = Y Jq^H):
//   block (p,q)[(i,l),(j,l')] = R( sum_t conj(X_lj) Y_il' ) S ,  sum_t conj(X_lj) Y_il'
//                             = sum_ab Jq_ja Jp_ib Th[(l,a),(b,l')]
//   block (p,p)[(i,l),(i,l')] = R( conj(Hp_ll') ),  Hp_ll' = sum_ab (Jq^H Jq)_ab Th[(l',b),(l,a)]
//   block (q,q)[(j,l),(j,l')] = R( conj(Hq_ll') ),  Hq_ll' = sum_ab (Jp^H Jp)_ab Th[(a,l),(b,l')]
// where R(z) = [[zr,-zi],[zi,zr]], S = diag(1,-1); parameter order per station is
// [Re J00, Im J00, Re J01, Im J01, Re J10, Im J10, Re J11, Im J11] (lmfit.c:90-97,460-467).
// See DESIGN.md for the derivation.
#include "internal.cuh"

// Hermitian 4x4 Gram tensor access from its packed form (k_coh_gram): index u = 2a+b
struct Gram {
  double d[4];
  double2 o[6];  // (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
  __device__ __forceinline__ double2 at(int u, int v) const {
    if (u == v) return make_double2(d[u], 0.0);
    int lo = u < v ? u : v, hi = u < v ? v : u;
    int idx = (lo == 0) ? (hi - 1) : (lo == 1 ? (hi + 1) : 5);
    double2 z = o[idx];
    return (u < v) ? z : make_double2(z.x, -z.y);
  }
};


// One baseline's contribution to J^T J.  The 16 2x2 sub-blocks of the (p,q) coupling (and their
// mirror images) are shared out over 16 threads (`sub`): each thread recomputes the few hundred flops
// (indices stay static, everything in registers) and writes only its own 8 entries, so the scattered
// stores of a baseline are spread over 16 threads and the grid has Nbase*16 threads instead of 13 CTAs.
__device__ __forceinline__ void assemble_baseline(const AssembleArgs &a, int p, int q, long long b,
                                                  int sub) {
  Gram G;
  {
    const double2 *Tb = reinterpret_cast<const double2 *>(a.T + b * 16);
    double2 t0 = Tb[0], t1 = Tb[1];
    G.d[0] = t0.x; G.d[1] = t0.y; G.d[2] = t1.x; G.d[3] = t1.y;
#pragma unroll
    for (int z = 0; z < 6; z++) G.o[z] = Tb[2 + z];
  }
  double2 Jp[4], Jq[4];
  load_jones(a.pblk, p, Jp);
  load_jones(a.pblk, q, Jq);
  const int ld = 8 * a.N;
  // cross block
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int l = 0; l < 2; l++)
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int lp = 0; lp < 2; lp++) {
          double2 z = make_double2(0.0, 0.0);
#pragma unroll
          for (int aa = 0; aa < 2; aa++)
#pragma unroll
            for (int bb = 0; bb < 2; bb++) {
              double2 jj = cmul(Jq[2 * j + aa], Jp[2 * i + bb]);
              cfma(z, jj, G.at(2 * l + aa, 2 * bb + lp));
            }
          if (sub != ((i * 2 + l) * 2 + j) * 2 + lp) continue;
          const int r0 = 8 * p + 2 * (2 * i + l);
          const int c0 = 8 * q + 2 * (2 * j + lp);
          // R(z) S = [[zr, zi],[zi, -zr]]
          a.JTJ[(long long)r0 * ld + c0] = z.x;
          a.JTJ[(long long)r0 * ld + c0 + 1] = z.y;
          a.JTJ[(long long)(r0 + 1) * ld + c0] = z.y;
          a.JTJ[(long long)(r0 + 1) * ld + c0 + 1] = -z.x;
          // transposed block (q,p)
          a.JTJ[(long long)c0 * ld + r0] = z.x;
          a.JTJ[(long long)(c0 + 1) * ld + r0] = z.y;
          a.JTJ[(long long)c0 * ld + r0 + 1] = z.y;
          a.JTJ[(long long)(c0 + 1) * ld + r0 + 1] = -z.x;
        }
  // diagonal contributions
  double2 Qq[4], Qp[4];
  mat_ahb(Jq, Jq, Qq);
  mat_ahb(Jp, Jp, Qp);
  double2 Hp[4], Hq[4];
#pragma unroll
  for (int l = 0; l < 2; l++)
#pragma unroll
    for (int lp = 0; lp < 2; lp++) {
      double2 hp = make_double2(0.0, 0.0), hq = make_double2(0.0, 0.0);
#pragma unroll
      for (int aa = 0; aa < 2; aa++)
#pragma unroll
        for (int bb = 0; bb < 2; bb++) {
          cfma(hp, Qq[2 * aa + bb], G.at(2 * lp + bb, 2 * l + aa));
          cfma(hq, Qp[2 * aa + bb], G.at(2 * aa + l, 2 * bb + lp));
        }
      Hp[2 * l + lp] = hp;
      Hq[2 * l + lp] = hq;
    }
  if (sub == 0) atomicAdd(a.Hst + 4 * p + 0, Hp[0].x);
  if (sub == 1) atomicAdd(a.Hst + 4 * p + 1, Hp[3].x);
  if (sub == 2) atomicAdd(a.Hst + 4 * p + 2, Hp[1].x);
  if (sub == 3) atomicAdd(a.Hst + 4 * p + 3, Hp[1].y);
  if (sub == 4) atomicAdd(a.Hst + 4 * q + 0, Hq[0].x);
  if (sub == 5) atomicAdd(a.Hst + 4 * q + 1, Hq[3].x);
  if (sub == 6) atomicAdd(a.Hst + 4 * q + 2, Hq[1].x);
  if (sub == 7) atomicAdd(a.Hst + 4 * q + 3, Hq[1].y);
}

// thread -> (baseline, sub-block): 16 consecutive threads share a baseline
__device__ __forceinline__ void assemble_lin(const AssembleArgs &a) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = gid >> 4;
  if (b >= a.Nbase) return;
  const short2 pq = a.blpq[b];
  assemble_baseline(a, pq.x, pq.y, b, (int)(gid & 15));
}

__global__ void __launch_bounds__(256)
k_assemble_offdiag(AssembleArgs a) {
  assemble_lin(a);
}

// batched over clusters (blockIdx.y): all normal matrices of a SAGE sweep in one launch
__global__ void __launch_bounds__(256)
k_assemble_offdiag_batched(BatchAssembleArgs b) {
  const int k = b.list[blockIdx.y];
  AssembleArgs a;
  a.T = b.T + (long long)b.tix[k] * b.Nbase * 16;
  a.pblk = b.pp + b.poff[k];
  a.JTJ = b.JTJ + (long long)blockIdx.y * 64 * b.N * b.N;
  a.Hst = b.Hst + (long long)blockIdx.y * 4 * b.N;
  a.tiles = b.tiles;
  a.blpq = b.blpq;
  a.N = b.N;
  a.Nbase = b.Nbase;
  assemble_lin(a);
}

__device__ __forceinline__ void assemble_diag_station(const double *__restrict__ Hst,
                                                      double *__restrict__ JTJ, int N, int s);

// diagonal 8x8 blocks from the station sums; one thread per (station, 4x4 sub-block i)
__global__ void k_assemble_diag(const double *__restrict__ Hst, double *__restrict__ JTJ, int N) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  assemble_diag_station(Hst, JTJ, N, s);
}
__global__ void k_assemble_diag_batched(const double *__restrict__ Hst, double *__restrict__ JTJ,
                                        int N) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  assemble_diag_station(Hst + (long long)blockIdx.y * 4 * N, JTJ + (long long)blockIdx.y * 64 * N * N,
                        N, s);
}
__device__ __forceinline__ void assemble_diag_station(const double *__restrict__ Hst,
                                                      double *__restrict__ JTJ, int N, int s) {
  const double h00 = Hst[4 * s], h11 = Hst[4 * s + 1], hr = Hst[4 * s + 2], hi = Hst[4 * s + 3];
  const int ld = 8 * N;
  for (int i = 0; i < 2; i++) {
    double blk[4][4] = {{h00, 0.0, hr, hi}, {0.0, h00, -hi, hr}, {hr, -hi, h11, 0.0},
                        {hi, hr, 0.0, h11}};
    for (int r = 0; r < 8; r++)
      for (int c = 0; c < 8; c++) {
        double v = 0.0;
        if ((r >> 2) == i && (c >> 2) == i) v = blk[r & 3][c & 3];
        if ((r >> 2) == i) JTJ[(long long)(8 * s + r) * ld + 8 * s + c] = v;
      }
  }
}

// per matrix of the batch: mu0 = tau * max_i (J^T J)_ii (clmfit.c:342-352) -> mu[y]; the diagonal is
// (h00 x4, h11 x4) per station, so the maximum runs over the station sums
__global__ void __launch_bounds__(128)
k_batch_mu0(const double *__restrict__ Hst, double *__restrict__ mu, int N, double tau) {
  __shared__ double sv[4];
  const double *H = Hst + (long long)blockIdx.x * 4 * N;
  double mx = 0.0;
  for (int s = threadIdx.x; s < N; s += blockDim.x) {
    const double a = H[4 * s], b = H[4 * s + 1];
    if (fabs(a) > fabs(mx)) mx = a;
    if (fabs(b) > fabs(mx)) mx = b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, mx, o);
    if (fabs(other) > fabs(mx)) mx = other;
  }
  if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 4; i++)
      if (fabs(sv[i]) > fabs(mx)) mx = sv[i];
    mu[blockIdx.x] = tau * mx;
  }
}

// A[y] = A0[y] + mu[y] I over a batch of n x n matrices
__global__ void k_batch_add_diag(const double *__restrict__ A0, double *__restrict__ A,
                                 const double *__restrict__ mu, int n) {
  const long long nn = (long long)n * n;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nn) {
    double v = A0[(long long)blockIdx.y * nn + i];
    if (i / n == i % n) v += mu[blockIdx.y];
    A[(long long)blockIdx.y * nn + i] = v;
  }
}

// A_ii += mu  (cudakernel_diagmu, mderiv.cu:936; my_daxpys on the diagonal, clmfit.c:363)
__global__ void k_copy_add_diag(const double *__restrict__ A0, double *__restrict__ A, int n,
                                double mu) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nn = (long long)n * n;
  if (i < nn) {
    double v = A0[i];
    if (i / n == i % n) v += mu;
    A[i] = v;
  }
}

__global__ void k_extract_diag(const double *__restrict__ A, double *__restrict__ dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = A[(long long)i * n + i];
}

// pnew = p + dp ; sc[0] = |dp|^2 ; sc[1] = dp . J^T e   (clmfit.c:440-449,487-497), one CTA
// zero (optional): accumulator of the trial pass that follows, cleared here instead of by a memset
__global__ void __launch_bounds__(512)
k_lm_step(const double *__restrict__ p, const double *__restrict__ Dp,
          const double *__restrict__ jte, double *__restrict__ pnew, double *__restrict__ sc,
          double *__restrict__ zero, int n) {
  __shared__ double s0[16], s1[16];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double dp = Dp[i];
    pnew[i] = p[i] + dp;
    if (zero) zero[i] = 0.0;
    a = fma(dp, dp, a);
    b = fma(dp, jte[i], b);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) {
    s0[threadIdx.x >> 5] = a;
    s1[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) {
      ta += s0[i];
      tb += s1[i];
    }
    sc[0] = ta;
    sc[1] = tb;
  }
}

// g <- g - y/2 - rho/2 (p - bz): J^T e of a pass -> (minus half) the gradient of the consensus-
// augmented cost ||e||^2 + y^T (p - bz) + rho/2 |p - bz|^2
__global__ void k_lm_aug_rhs(double *__restrict__ g, const double *__restrict__ p,
                             const double *__restrict__ y, const double *__restrict__ bz, double rho,
                             int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g[i] = g[i] - 0.5 * y[i] - 0.5 * rho * (p[i] - bz[i]);
}

extern "C" {
void db_launch_lm_aug_rhs(double *jte, const double *p, const double *y, const double *bz,
                          double rho, int n, cudaStream_t st) {
  k_lm_aug_rhs<<<(n + 255) / 256, 256, 0, st>>>(jte, p, y, bz, rho, n);
}
void db_launch_lm_step(const double *p, const double *Dp, const double *jte, double *pnew,
                       double *sc, double *zero, int n, cudaStream_t st) {
  k_lm_step<<<1, 512, 0, st>>>(p, Dp, jte, pnew, sc, zero, n);
}
void db_launch_extract_diag(const double *A, double *dst, int n, cudaStream_t st) {
  k_extract_diag<<<(n + 127) / 128, 128, 0, st>>>(A, dst, n);
}
void db_launch_assemble(const AssembleArgs *a, int ntile, cudaStream_t st) {
  (void)ntile;
  k_assemble_offdiag<<<(a->Nbase * 16 + 255) / 256, 256, 0, st>>>(*a);
  k_assemble_diag<<<(a->N + 63) / 64, 64, 0, st>>>(a->Hst, a->JTJ, a->N);
}
void db_launch_assemble_batched(const BatchAssembleArgs *b, int ntile, int nb, double tau,
                                double *mu, double *Afac, cudaStream_t st) {
  (void)ntile;
  dim3 g1((b->Nbase * 16 + 255) / 256, nb);
  k_assemble_offdiag_batched<<<g1, 256, 0, st>>>(*b);
  dim3 g2((b->N + 63) / 64, nb);
  k_assemble_diag_batched<<<g2, 64, 0, st>>>(b->Hst, b->JTJ, b->N);
  k_batch_mu0<<<nb, 128, 0, st>>>(b->Hst, mu, b->N, tau);
  if (!Afac) return;  // the cluster Cholesky reads J^T J and mu itself
  const int n = 8 * b->N;
  const long long nn = (long long)n * n;
  dim3 g3((unsigned)((nn + 255) / 256), nb);
  k_batch_add_diag<<<g3, 256, 0, st>>>(b->JTJ, Afac, mu, n);
}
void db_launch_copy_add_diag(const double *A0, double *A, int n, double mu, cudaStream_t st) {
  long long nn = (long long)n * n;
  k_copy_add_diag<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(A0, A, n, mu);
}
}
