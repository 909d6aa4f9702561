// Internal: packed sky model for the device-side coherency prediction.
#pragma once
#include "internal.cuh"

#define STYPE_POINT_ 0
#define STYPE_GAUSSIAN_ 1
#define STYPE_DISK_ 2
#define STYPE_RING_ 3
#define STYPE_SHAPELET_ 4

// one source, 224 bytes (multiple of 16: TMA bulk-copy granularity)
struct DevSource {
  double ll, mm, nn, sI, sQ, sU, sV, stype;
  double eX, eY, eP, cxi, sxi, cphi, sphi, use_projection;
  double sI0, sQ0, sU0, sV0, f0, spec_idx, spec_idx1, spec_idx2;
  double sh_n0, sh_beta, sh_off, pad_;  // shapelets: order, scale, first coefficient in CohArgs::modes
};

#define COH_SEG_MAX 96  // sources staged per bulk copy (2 x 96 x 224 B = 42 KB of smem)
#define COH_SHAPELET_MAX_N0 32  // largest shapelet order the device kernel takes

// a run of <= COH_SEG_MAX sources of one cluster
struct CohSegment {
  int first;    // index into the packed source array
  int count;
  int cluster;
  int last;     // 1 if this run closes its cluster
};

struct CohArgs {
  const double *u, *v, *w;   // [R] seconds
  const DevSource *src;
  const double *modes;       // shapelet coefficients of all sources, back to back (may be null)
  const CohSegment *segs;
  int nseg;
  const double *freqs;       // device, [Nchan]
  int Nchan;
  double fdelta2;            // half channel width used for the |sinc| smearing
  double uvmin, uvmax;
  long long R;
  double2 *coh;              // MODE 0 out: planar [M][4][R]
  unsigned char *flag;       // MODE 0: uv-cut flags (may be null)
  double2 *xout;             // MODE 1, 2 in/out: [Nchan][R][4]
  // MODE 2 (full-resolution residual with solutions, residual.c:681-938)
  const int *sta1, *sta2;        // [R] stations of every row
  const double *p;               // Jones solutions (layout of pp)
  const int *clus_nchunk;        // [M]
  const int *clus_chunk0;        // [M] first entry of cluster k in chunk_poff
  const int *chunk_poff;         // [Mt]
  const unsigned char *clus_sub; // [M] 1: subtract this cluster (id >= 0)
  const double *pinv;            // inverse Jones of the correction cluster [nchunk][N][8], or null
  int pinv_nchunk, N;
};

extern "C" {
void db_launch_coherencies(const CohArgs *a, cudaStream_t st);
void db_launch_predict_multifreq(const CohArgs *a, cudaStream_t st);
void db_launch_residual_multifreq(const CohArgs *a, cudaStream_t st);
}
