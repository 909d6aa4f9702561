// Triangular solves L L^T x = b for systems beyond the cluster kernels (8N > 512; 512 stations: n = 4096),
// on a factor cuSOLVER's dpotrf left (column-major lower, ld).  Replaces cusolverDnDpotrs, whose two
// trsv kernels take 0.68 ms for 2 x 67 MB of traffic: a dependency chain of n/64 block steps, bound by
// latency, not by bytes.
//
// Scheme (the same dataflow as k_tri_solve, with global flags instead of distributed shared memory):
// 64 x 64 blocks, one CTA per block row (forward) / block column (backward), all co-resident.  CTA i
// accumulates b_i - sum_{k<i} L_ik y_k as the y_k arrive (it polls one release/acquire flag per block,
// the next off-chain block already in registers), keeps the one block that sits on the chain
// (L_{i,i-1}) and the inverse of its diagonal block in shared memory, and publishes y_i = L_ii^-1 (...)
// with a release store.  The chain per block is: flag -> 64 x 64 product from shared memory -> 64 x 64
// product with the inverse -> flag.  The diagonal inverses come from a small kernel of their own (one
// CTA per block, one thread per column), once per factor.
#include "internal.cuh"

#define BT 64          // block size
#define BT_THREADS 256  // 4 quarters x 64
#define BT_PENDING 0x7ff8dead0badbeefull  // quiet NaN with a payload no arithmetic produces

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// inverse of every 64 x 64 diagonal block of L: Linv[blk][c*64 + r] (column-major), thread c solves
// L x = e_c by forward substitution with the block in shared memory (broadcast reads)
__global__ void __launch_bounds__(BT)
k_bigtri_diag_inv(const double *__restrict__ L, int ld, double *__restrict__ Linv) {
  __shared__ double Ls[BT * BT];
  const int blk = blockIdx.x, c = threadIdx.x;
  const double *src = L + (size_t)(blk * BT) * ld + blk * BT;
  for (int e = threadIdx.x; e < BT * BT; e += BT) {
    const int cc = e / BT, r = e % BT;
    Ls[cc * BT + r] = (r >= cc) ? src[(size_t)cc * ld + r] : 0.0;
  }
  __syncthreads();
  double x[BT];
#pragma unroll
  for (int r = 0; r < BT; r++) x[r] = 0.0;
#pragma unroll
  for (int r = 0; r < BT; r++) {
    double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < BT; k++)
      if (k < r) s = fma(-Ls[k * BT + r], x[k], s);
    x[r] = (r >= c) ? s / Ls[r * BT + r] : 0.0;
  }
  double *dst = Linv + (size_t)blk * BT * BT + (size_t)c * BT;
#pragma unroll
  for (int r = 0; r < BT; r++) dst[r] = x[r];
}

struct BigTriArgs {
  const double *L;
  const double *Linv;   // [nb][64*64] column-major inverses of the diagonal blocks
  const double *rhs;    // forward: b ; backward: y
  double *out;          // forward: y ; backward: x
  unsigned *flags;      // [nb] arrival flags of this direction
  unsigned epoch;       // value that marks "published" for this solve
  int n, ld, nb;
};

// FWD: y_i = Linv_ii (b_i - sum_{k<i} L_ik y_k), CTA i = block row i, k ascending.
// !FWD: x_j = Linv_jj^T (y_j - sum_{i>j} L_ij^T x_i), CTA j = block column nb-1-blockIdx.x, i descending.
template <bool FWD>
__global__ void __launch_bounds__(BT_THREADS)
k_bigtri(BigTriArgs a) {
  extern __shared__ __align__(16) double sm[];
  double *Lc = sm;                 // the on-chain block: FWD L_{i,i-1} as stored, !FWD L_{j+1,j} transposed
  double *Li = Lc + BT * BT;       // FWD Linv_ii as stored, !FWD Linv_jj transposed
  double *vec = Li + BT * BT;      // [64] incoming block of the solution
  double *part = vec + BT;         // [4][64]
  const int tid = threadIdx.x, r = tid & (BT - 1), q = tid >> 6;
  const int me = FWD ? (int)blockIdx.x : a.nb - 1 - (int)blockIdx.x;
  const int nprev = FWD ? me : a.nb - 1 - me;  // blocks this CTA consumes
  // stage the on-chain block and the diagonal inverse
  {
    const double *inv = a.Linv + (size_t)me * BT * BT;
    for (int e = tid; e < BT * BT; e += BT_THREADS) {
      const int cc = e / BT, rr = e % BT;
      if (FWD) Li[cc * BT + rr] = inv[e];
      else Li[rr * BT + cc] = inv[e];  // transposed: Li[c + r*64] = Linv[r][c]
    }
    if (nprev > 0) {
      const int bi = FWD ? me : me + 1, bk = FWD ? me - 1 : me;  // block (bi, bk) of L
      const double *src = a.L + (size_t)(bk * BT) * a.ld + bi * BT;
      for (int e = tid; e < BT * BT; e += BT_THREADS) {
        const int cc = e / BT, rr = e % BT;
        const double v = src[(size_t)cc * a.ld + rr];
        if (FWD) Lc[cc * BT + rr] = v;
        else Lc[rr * BT + cc] = v;  // transposed: Lc[c + r*64] = L[r][c]
      }
    }
  }
  double acc = 0.0;
  // off-chain blocks, next one prefetched into registers while the flag of the current one is awaited.
  // FWD : thread (r, q) holds L[(me*64 + r), (k*64 + 16q .. 16q+15)]          (rows across threads)
  // !FWD: thread (c=r, q) holds L[(i*64 + 16q .. 16q+15), (me*64 + c)]        (a run of 16 rows)
  double nx[16];
  auto fetch = [&](int step) {
    const int other = FWD ? step : a.nb - 1 - step;  // k ascending / i descending
    if (FWD) {
      const double *src = a.L + (size_t)(other * BT + 16 * q) * a.ld + me * BT + r;
#pragma unroll
      for (int c = 0; c < 16; c++) nx[c] = __ldcg(src + (size_t)c * a.ld);
    } else {
      const double *src = a.L + (size_t)(me * BT + r) * a.ld + other * BT + 16 * q;
#pragma unroll
      for (int c = 0; c < 16; c++) nx[c] = __ldcg(src + c);
    }
  };
  const int noff = nprev > 0 ? nprev - 1 : 0;  // all but the on-chain one
  if (noff > 0) fetch(0);
  __syncthreads();
  for (int step = 0; step < nprev; step++) {
    const int other = FWD ? step : a.nb - 1 - step;
    // the block of the solution is its own arrival flag: the slots start as a NaN pattern no computation
    // produces, 64-bit stores are single-copy atomic, every thread watches one element (one L2 round
    // trip per step instead of flag + data)
    if (tid < BT) {
      const unsigned long long *src = reinterpret_cast<const unsigned long long *>(a.out) +
                                      (size_t)other * BT + tid;
      unsigned long long bits;
      do {
        asm volatile("ld.relaxed.gpu.u64 %0, [%1];" : "=l"(bits) : "l"(src) : "memory");
      } while (bits == BT_PENDING);
      vec[tid] = __longlong_as_double((long long)bits);
    }
    __syncthreads();
    if (step < noff) {
      double cur[16];
#pragma unroll
      for (int c = 0; c < 16; c++) cur[c] = nx[c];
      if (step + 1 < noff) fetch(step + 1);
#pragma unroll
      for (int c = 0; c < 16; c++) acc = fma(-cur[c], vec[16 * q + c], acc);
    } else {
      // the on-chain block from shared memory: both layouts read Lc[(16q + c)*64 + r]
#pragma unroll
      for (int c = 0; c < 16; c++) acc = fma(-Lc[(16 * q + c) * BT + r], vec[16 * q + c], acc);
    }
    __syncthreads();  // vec is rewritten by the next step
  }
  part[q * BT + r] = acc;
  __syncthreads();
  if (tid < BT)
    vec[tid] = a.rhs[(size_t)me * BT + tid] + part[tid] + part[BT + tid] + part[2 * BT + tid] +
               part[3 * BT + tid];
  __syncthreads();
  // out_me = Linv (FWD) / Linv^T (!FWD) times vec: both layouts read Li[(16q + c)*64 + r]
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < 16; c++) s = fma(Li[(16 * q + c) * BT + r], vec[16 * q + c], s);
  __syncthreads();
  part[q * BT + r] = s;
  __syncthreads();
  if (tid < BT) {
    const double v = part[tid] + part[BT + tid] + part[2 * BT + tid] + part[3 * BT + tid];
    asm volatile("st.relaxed.gpu.f64 [%0], %1;" ::"l"(a.out + (size_t)me * BT + tid), "d"(v) : "memory");
  }
}

// every slot of y and x pending (before the forward kernel of a solve)
__global__ void k_bigtri_arm(unsigned long long *y, unsigned long long *x, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    y[i] = BT_PENDING;
    x[i] = BT_PENDING;
  }
}

extern "C" {
// 1 if this size is handled (multiple of 64, all CTAs co-resident)
int db_bigtri_available(int n) {
  if (getenv("DIRAC_B200_NO_BIGTRI")) return 0;
  return n > 512 && (n % BT) == 0 && n / BT <= db_sm_count();
}
size_t db_bigtri_ws_doubles(int n) { return (size_t)(n / BT) * BT * BT + (size_t)n + (size_t)(n / BT) + 16; }

// L L^T x = b.  ws: db_bigtri_ws_doubles(n) doubles (diagonal inverses | y | flags); epoch: a value that
// differs from call to call (the flags are never reset).  invert: recompute the diagonal inverses (new
// factor).
void db_launch_bigtri_solve(const double *L, int ld, int n, const double *b, double *x, double *ws,
                            unsigned epoch, int invert, cudaStream_t st) {
  const int nb = n / BT;
  double *Linv = ws, *y = ws + (size_t)nb * BT * BT;
  unsigned *flags = reinterpret_cast<unsigned *>(y + n);
  static bool configured = false;
  const size_t smem = sizeof(double) * (2 * BT * BT + BT + 4 * BT);
  if (!configured) {
    DB_CHECK(cudaFuncSetAttribute(k_bigtri<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DB_CHECK(cudaFuncSetAttribute(k_bigtri<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  if (invert) k_bigtri_diag_inv<<<nb, BT, 0, st>>>(L, ld, Linv);
  k_bigtri_arm<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<unsigned long long *>(y),
                                               reinterpret_cast<unsigned long long *>(x), n);
  BigTriArgs a;
  a.L = L; a.Linv = Linv; a.n = n; a.ld = ld; a.nb = nb;
  a.rhs = b; a.out = y; a.flags = flags; a.epoch = 2 * epoch + 1;
  k_bigtri<true><<<nb, BT_THREADS, smem, st>>>(a);
  a.rhs = y; a.out = x; a.flags = flags + nb; a.epoch = 2 * epoch + 2;
  k_bigtri<false><<<nb, BT_THREADS, smem, st>>>(a);
}

// test / tuning hook: x = (L L^T)^-1 b from host buffers (L column-major lower, ld = n); reps > 0
// additionally times `reps` back-to-back solves on the resident factor (us per solve in *us).
// returns -1 when the size is not handled
int dirac_b200_bigtri_solve(int n, const double *L, const double *b, double *x, int reps, double *us) {
  if (!db_bigtri_available(n)) return -1;
  double *dL, *db, *dx, *ws;
  const size_t nd = db_bigtri_ws_doubles(n);
  DB_CHECK(cudaMalloc((void **)&dL, sizeof(double) * (size_t)n * n));
  DB_CHECK(cudaMalloc((void **)&db, sizeof(double) * n));
  DB_CHECK(cudaMalloc((void **)&dx, sizeof(double) * n));
  DB_CHECK(cudaMalloc((void **)&ws, sizeof(double) * nd));
  DB_CHECK(cudaMemset(ws, 0, sizeof(double) * nd));
  DB_CHECK(cudaMemcpy(dL, L, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
  DB_CHECK(cudaMemcpy(db, b, sizeof(double) * n, cudaMemcpyHostToDevice));
  unsigned epoch = 0;
  db_launch_bigtri_solve(dL, n, n, db, dx, ws, ++epoch, 1, 0);
  DB_CHECK(cudaDeviceSynchronize());
  DB_CHECK(cudaMemcpy(x, dx, sizeof(double) * n, cudaMemcpyDeviceToHost));
  if (reps > 0 && us) {
    cudaEvent_t e0, e1;
    DB_CHECK(cudaEventCreate(&e0));
    DB_CHECK(cudaEventCreate(&e1));
    DB_CHECK(cudaEventRecord(e0, 0));
    for (int i = 0; i < reps; i++) db_launch_bigtri_solve(dL, n, n, db, dx, ws, ++epoch, 0, 0);
    DB_CHECK(cudaEventRecord(e1, 0));
    DB_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    DB_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *us = 1e3 * ms / reps;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  DB_CHECK(cudaGetLastError());
  cudaFree(dL); cudaFree(db); cudaFree(dx); cudaFree(ws);
  return 0;
}
}
