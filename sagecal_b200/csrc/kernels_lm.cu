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


__global__ void __launch_bounds__(TILE_THREADS)
k_assemble_offdiag(AssembleArgs a) {
  const TileDesc td = a.tiles[blockIdx.x];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = td.pb * TILE_P + w;
  const int q = td.qb * TILE_Q + lane;
  if (!((q > p) && (q < a.N))) return;
  const long long b = baseline_index(p, q, a.N);
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
  atomicAdd(a.Hst + 4 * p + 0, Hp[0].x);
  atomicAdd(a.Hst + 4 * p + 1, Hp[3].x);
  atomicAdd(a.Hst + 4 * p + 2, Hp[1].x);
  atomicAdd(a.Hst + 4 * p + 3, Hp[1].y);
  atomicAdd(a.Hst + 4 * q + 0, Hq[0].x);
  atomicAdd(a.Hst + 4 * q + 1, Hq[3].x);
  atomicAdd(a.Hst + 4 * q + 2, Hq[1].x);
  atomicAdd(a.Hst + 4 * q + 3, Hq[1].y);
}

// diagonal 8x8 blocks from the station sums; one thread per (station, 4x4 sub-block i)
__global__ void k_assemble_diag(const double *__restrict__ Hst, double *__restrict__ JTJ, int N) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
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

extern "C" {
void db_launch_extract_diag(const double *A, double *dst, int n, cudaStream_t st) {
  k_extract_diag<<<(n + 127) / 128, 128, 0, st>>>(A, dst, n);
}
void db_launch_assemble(const AssembleArgs *a, int ntile, cudaStream_t st) {
  k_assemble_offdiag<<<ntile, TILE_THREADS, 0, st>>>(*a);
  k_assemble_diag<<<(a->N + 63) / 64, 64, 0, st>>>(a->Hst, a->JTJ, a->N);
}
void db_launch_copy_add_diag(const double *A0, double *A, int n, double mu, cudaStream_t st) {
  long long nn = (long long)n * n;
  k_copy_add_diag<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(A0, A, n, mu);
}
}
