// Host side of the device coherency prediction: packs clus_source_t into the flat device sky,
// and implements precalculate_coherencies / predict_visibilities_multifreq of the Dirac radio API.
#include <math.h>
#include <string.h>
#include <vector>

#include "../../include/dirac_b200.h"
#include "coh.h"
#include "problem.h"

struct SkyDev {
  DevSource *src;
  CohSegment *segs;
  int nseg;
};

static void sky_upload(const clus_source_t *carr, int M, SkyDev *sky, cudaStream_t st) {
  std::vector<DevSource> src;
  std::vector<CohSegment> segs;
  for (int k = 0; k < M; k++) {
    const clus_source_t &c = carr[k];
    const int first = (int)src.size();
    for (int s = 0; s < c.N; s++) {
      DevSource d;
      memset(&d, 0, sizeof(d));
      d.ll = c.ll[s]; d.mm = c.mm[s]; d.nn = c.nn[s];
      d.sI = c.sI[s]; d.sQ = c.sQ[s]; d.sU = c.sU[s]; d.sV = c.sV[s];
      d.stype = (double)c.stype[s];
      if (c.stype[s] == STYPE_SHAPELET) {
        fprintf(stderr, "dirac_b200: shapelet sources are not supported by the device "
                        "coherency kernel (cluster %d source %d)\n", k, s);
        exit(1);
      }
      if (c.stype[s] == STYPE_GAUSSIAN && c.ex && c.ex[s]) {
        const exinfo_gaussian *g = (const exinfo_gaussian *)c.ex[s];
        d.eX = g->eX; d.eY = g->eY; d.eP = g->eP; d.cxi = g->cxi; d.sxi = g->sxi;
        d.cphi = g->cphi; d.sphi = g->sphi; d.use_projection = (double)g->use_projection;
      } else if ((c.stype[s] == STYPE_DISK || c.stype[s] == STYPE_RING) && c.ex && c.ex[s]) {
        // exinfo_disk / exinfo_ring: { eX; cxi, sxi, cphi, sphi; use_projection }
        const double *g = (const double *)c.ex[s];
        d.eX = g[0]; d.cxi = g[1]; d.sxi = g[2]; d.cphi = g[3]; d.sphi = g[4];
        d.use_projection = 1.0;
      }
      d.sI0 = c.sI0 ? c.sI0[s] : c.sI[s];
      d.sQ0 = c.sQ0 ? c.sQ0[s] : c.sQ[s];
      d.sU0 = c.sU0 ? c.sU0[s] : c.sU[s];
      d.sV0 = c.sV0 ? c.sV0[s] : c.sV[s];
      d.f0 = c.f0 ? c.f0[s] : 1.0;
      d.spec_idx = c.spec_idx ? c.spec_idx[s] : 0.0;
      d.spec_idx1 = c.spec_idx1 ? c.spec_idx1[s] : 0.0;
      d.spec_idx2 = c.spec_idx2 ? c.spec_idx2[s] : 0.0;
      src.push_back(d);
    }
    int done = 0;
    do {  // an empty cluster still yields one (empty, closing) segment
      CohSegment sg;
      sg.first = first + done;
      sg.count = c.N - done;
      if (sg.count > COH_SEG_MAX) sg.count = COH_SEG_MAX;
      sg.cluster = k;
      done += sg.count;
      sg.last = (done >= c.N) ? 1 : 0;
      segs.push_back(sg);
    } while (done < c.N);
  }
  if (src.empty()) src.resize(1);
  DB_CHECK(cudaMalloc((void **)&sky->src, sizeof(DevSource) * src.size()));
  DB_CHECK(cudaMalloc((void **)&sky->segs, sizeof(CohSegment) * segs.size()));
  DB_CHECK(cudaMemcpyAsync(sky->src, src.data(), sizeof(DevSource) * src.size(),
                           cudaMemcpyHostToDevice, st));
  DB_CHECK(cudaMemcpyAsync(sky->segs, segs.data(), sizeof(CohSegment) * segs.size(),
                           cudaMemcpyHostToDevice, st));
  db_stream_sync(st);  // the vectors go out of scope
  sky->nseg = (int)segs.size();
}

static void sky_free(SkyDev *sky) {
  cudaFree(sky->src);
  cudaFree(sky->segs);
}

static double *upload_doubles(const double *h, size_t n, cudaStream_t st) {
  double *d = nullptr;
  DB_CHECK(cudaMalloc((void **)&d, sizeof(double) * (n ? n : 1)));
  DB_CHECK(cudaMemcpyAsync(d, h, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  return d;
}

extern "C" void dirac_b200_precalculate(dirac_b200_problem *pr, const double *u, const double *v,
                                        const double *w, const clus_source_t *carr, double freq0,
                                        double fdelta, double uvmin, double uvmax,
                                        baseline_t *barr) {
  DevProblem &d = pr->d;
  SkyDev sky;
  sky_upload(carr, d.M, &sky, d.stream);
  double *du = upload_doubles(u, d.R, d.stream);
  double *dv = upload_doubles(v, d.R, d.stream);
  double *dw = upload_doubles(w, d.R, d.stream);
  double *df = upload_doubles(&freq0, 1, d.stream);
  CohArgs a;
  memset(&a, 0, sizeof(a));
  a.u = du; a.v = dv; a.w = dw; a.src = sky.src; a.segs = sky.segs; a.nseg = sky.nseg;
  a.freqs = df; a.Nchan = 1; a.fdelta2 = fdelta * 0.5; a.uvmin = uvmin; a.uvmax = uvmax;
  a.R = d.R; a.coh = d.coh; a.flag = d.flag; a.xout = nullptr;
  db_launch_coherencies(&a, d.stream);
  db_count_launch(1);
  if (barr) {
    std::vector<unsigned char> hf(d.R);
    DB_CHECK(cudaMemcpyAsync(hf.data(), d.flag, d.R, cudaMemcpyDeviceToHost, d.stream));
    db_stream_sync(d.stream);
    for (long long r = 0; r < d.R; r++) barr[r].flag = hf[r];
  }
  db_stream_sync(d.stream);
  DB_CHECK(cudaGetLastError());
  cudaFree(du); cudaFree(dv); cudaFree(dw); cudaFree(df);
  sky_free(&sky);
  // the Gram tensors cached for LM belong to the old coherencies
  if (pr->lm.ready) memset(pr->lm.T_valid, 0, d.Mt);
}

// Dirac_radio.h:209 — here Nbase is already Nbase*tilesz (predict.c:503-578); rows need not be in
// canonical order for this call (no station indexing), only flags are read/written.
extern "C" int precalculate_coherencies(double *u, double *v, double *w, double *x, int N,
                                        int Nbase, baseline_t *barr, clus_source_t *carr, int M,
                                        double freq0, double fdelta, double tdelta, double dec0,
                                        double uvmin, double uvmax, int Nt) {
  (void)N; (void)tdelta; (void)dec0; (void)Nt;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "dirac_b200: no CUDA device available. This library has no CPU fallback.\n");
    exit(1);
  }
  const long long R = Nbase;
  cudaStream_t st;
  DB_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  SkyDev sky;
  sky_upload(carr, M, &sky, st);
  double *du = upload_doubles(u, R, st), *dv = upload_doubles(v, R, st);
  double *dw = upload_doubles(w, R, st), *df = upload_doubles(&freq0, 1, st);
  std::vector<unsigned char> hf(R);
  for (long long r = 0; r < R; r++) hf[r] = barr[r].flag;
  unsigned char *dflag = nullptr;
  DB_CHECK(cudaMalloc((void **)&dflag, R + 16));
  DB_CHECK(cudaMemcpyAsync(dflag, hf.data(), R, cudaMemcpyHostToDevice, st));
  double2 *dcoh = nullptr;
  DB_CHECK(cudaMalloc((void **)&dcoh, sizeof(double2) * (size_t)M * 4 * R));
  CohArgs a;
  memset(&a, 0, sizeof(a));
  a.u = du; a.v = dv; a.w = dw; a.src = sky.src; a.segs = sky.segs; a.nseg = sky.nseg;
  a.freqs = df; a.Nchan = 1; a.fdelta2 = fdelta * 0.5; a.uvmin = uvmin; a.uvmax = uvmax;
  a.R = R; a.coh = dcoh; a.flag = dflag; a.xout = nullptr;
  db_launch_coherencies(&a, st);
  db_count_launch(1);
  // planar -> API layout in row blocks, D2H
  long long rows_per = (128ll << 20) / ((long long)M * 64);
  if (rows_per < 1) rows_per = 1;
  if (rows_per > R) rows_per = R;
  double2 *stage = nullptr;
  DB_CHECK(cudaMalloc((void **)&stage, sizeof(double2) * (size_t)rows_per * M * 4));
  for (long long r0 = 0; r0 < R; r0 += rows_per) {
    int nr = (int)((R - r0 < rows_per) ? (R - r0) : rows_per);
    db_launch_coh_from_planar(dcoh, stage, r0, nr, M, R, st);
    db_count_launch(1);
    DB_CHECK(cudaMemcpyAsync(x + (size_t)r0 * M * 8, stage, (size_t)nr * M * 64,
                             cudaMemcpyDeviceToHost, st));
  }
  DB_CHECK(cudaMemcpyAsync(hf.data(), dflag, R, cudaMemcpyDeviceToHost, st));
  db_stream_sync(st);
  DB_CHECK(cudaGetLastError());
  for (long long r = 0; r < R; r++) barr[r].flag = hf[r];
  cudaFree(stage); cudaFree(dcoh); cudaFree(dflag);
  cudaFree(du); cudaFree(dv); cudaFree(dw); cudaFree(df);
  sky_free(&sky);
  cudaStreamDestroy(st);
  return 0;
}

// Dirac_radio.h:659 (residual.c:1257-1340): x[chan][row][8] += sum over clusters; add_to_data ==
// SIMUL_ONLY (1, Dirac_radio.h:78) clears x first, every other value accumulates onto the input
// (the thread function only ever adds, residual.c:1238-1245).  No Jones, no flags.
extern "C" int predict_visibilities_multifreq(double *u, double *v, double *w, double *x, int N,
                                              int Nbase, int tilesz, baseline_t *barr,
                                              clus_source_t *carr, int M, double *freqs, int Nchan,
                                              double fdelta, double tdelta, double dec0, int Nt,
                                              int add_to_data) {
  (void)N; (void)barr; (void)tdelta; (void)dec0; (void)Nt;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "dirac_b200: no CUDA device available. This library has no CPU fallback.\n");
    exit(1);
  }
  const long long R = (long long)Nbase * tilesz;
  cudaStream_t st;
  DB_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  SkyDev sky;
  sky_upload(carr, M, &sky, st);
  double *du = upload_doubles(u, R, st), *dv = upload_doubles(v, R, st);
  double *dw = upload_doubles(w, R, st), *df = upload_doubles(freqs, Nchan, st);
  double2 *dx = nullptr;
  const size_t nx = (size_t)Nchan * R * 4;
  DB_CHECK(cudaMalloc((void **)&dx, sizeof(double2) * nx));
  if (add_to_data == 1) {  // SIMUL_ONLY
    DB_CHECK(cudaMemsetAsync(dx, 0, sizeof(double2) * nx, st));
  } else {
    DB_CHECK(cudaMemcpyAsync(dx, x, sizeof(double2) * nx, cudaMemcpyHostToDevice, st));
  }
  CohArgs a;
  memset(&a, 0, sizeof(a));
  a.u = du; a.v = dv; a.w = dw; a.src = sky.src; a.segs = sky.segs; a.nseg = sky.nseg;
  a.freqs = df; a.Nchan = Nchan; a.fdelta2 = (fdelta / (double)Nchan) * 0.5;
  a.R = R; a.xout = dx;
  db_launch_predict_multifreq(&a, st);
  db_count_launch(1);
  DB_CHECK(cudaMemcpyAsync(x, dx, sizeof(double2) * nx, cudaMemcpyDeviceToHost, st));
  db_stream_sync(st);
  DB_CHECK(cudaGetLastError());
  cudaFree(dx); cudaFree(du); cudaFree(dv); cudaFree(dw); cudaFree(df);
  sky_free(&sky);
  cudaStreamDestroy(st);
  return 0;
}
