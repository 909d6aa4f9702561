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
  double *modes;
  CohSegment *segs;
  int nseg;
  int nsrc;
};

static void sky_upload(const clus_source_t *carr, int M, SkyDev *sky, cudaStream_t st) {
  std::vector<DevSource> src;
  std::vector<double> modes;
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
      d.ra = c.ra ? c.ra[s] : 0.0;
      d.dec = c.dec ? c.dec[s] : 0.0;
      if (c.stype[s] == STYPE_SHAPELET && c.ex && c.ex[s]) {
        const exinfo_shapelet *g = (const exinfo_shapelet *)c.ex[s];
        if (g->n0 < 1 || g->n0 > COH_SHAPELET_MAX_N0) {
          fprintf(stderr, "dirac_b200: shapelet order %d of cluster %d source %d is outside 1..%d\n",
                  g->n0, k, s, COH_SHAPELET_MAX_N0);
          exit(1);
        }
        d.eX = g->eX; d.eY = g->eY; d.eP = g->eP; d.cxi = g->cxi; d.sxi = g->sxi;
        d.cphi = g->cphi; d.sphi = g->sphi; d.use_projection = (double)g->use_projection;
        d.sh_n0 = (double)g->n0; d.sh_beta = g->beta; d.sh_off = (double)modes.size();
        modes.insert(modes.end(), g->modes, g->modes + (size_t)g->n0 * g->n0);
      } else if (c.stype[s] == STYPE_SHAPELET) {
        fprintf(stderr, "dirac_b200: shapelet source without exinfo (cluster %d source %d)\n", k, s);
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
  if (modes.empty()) modes.resize(2, 0.0);
  DB_CHECK(cudaMalloc((void **)&sky->modes, sizeof(double) * modes.size()));
  DB_CHECK(cudaMemcpyAsync(sky->modes, modes.data(), sizeof(double) * modes.size(),
                           cudaMemcpyHostToDevice, st));
  DB_CHECK(cudaMalloc((void **)&sky->src, sizeof(DevSource) * src.size()));
  DB_CHECK(cudaMalloc((void **)&sky->segs, sizeof(CohSegment) * segs.size()));
  DB_CHECK(cudaMemcpyAsync(sky->src, src.data(), sizeof(DevSource) * src.size(),
                           cudaMemcpyHostToDevice, st));
  DB_CHECK(cudaMemcpyAsync(sky->segs, segs.data(), sizeof(CohSegment) * segs.size(),
                           cudaMemcpyHostToDevice, st));
  db_stream_sync(st);  // the vectors go out of scope
  sky->nseg = (int)segs.size();
  sky->nsrc = (int)src.size();
}

static void sky_free(SkyDev *sky) {
  cudaFree(sky->src);
  cudaFree(sky->modes);
  cudaFree(sky->segs);
}

static double *upload_doubles(const double *h, size_t n, cudaStream_t st) {
  double *d = nullptr;
  DB_CHECK(cudaMalloc((void **)&d, sizeof(double) * (n ? n : 1)));
  DB_CHECK(cudaMemcpyAsync(d, h, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  return d;
}


// ---- station beams (precalculate_coherencies_withbeam & co, predict_withbeam.c) -----------------------
struct BeamSpec {
  int bf_type;
  double b_ra0, b_dec0, ph_ra0, ph_dec0, ph_freq0;
  const double *longitude, *latitude, *time_utc;
  int tilesz;
  const int *Nelem;
  double **xx, **yy, **zz;
  const elementcoeff *ecoeff;
  int doBeam;
};
struct BeamDev {
  double *af;
  double2 *E;
  std::vector<void *> owned;
};
// builds the per (timeslot, channel, source, station) tables on the device and points the coherency
// kernel at them.  doBeam: Dirac_common.h:120-151
static void beam_prepare(const BeamSpec *b, const SkyDev &sky, int N, int Nbase_slot,
                         const double *freqs_host, int Nf, const double *freqs_dev,
                         const baseline_t *barr, long long R, cudaStream_t st, CohArgs *a,
                         BeamDev *bd) {
  bd->af = nullptr;
  bd->E = nullptr;
  if (!b || b->doBeam == DOBEAM_NONE) return;
  const bool wide = b->doBeam == DOBEAM_ARRAY_WB || b->doBeam == DOBEAM_FULL_WB ||
                    b->doBeam == DOBEAM_ELEMENT_WB;
  const bool do_array = b->doBeam == DOBEAM_ARRAY || b->doBeam == DOBEAM_FULL ||
                        b->doBeam == DOBEAM_ARRAY_WB || b->doBeam == DOBEAM_FULL_WB;
  const bool do_elem = b->doBeam == DOBEAM_ELEMENT || b->doBeam == DOBEAM_FULL ||
                       b->doBeam == DOBEAM_ELEMENT_WB || b->doBeam == DOBEAM_FULL_WB;
  if (!do_array && !do_elem) {
    fprintf(stderr, "dirac_b200: beam mode %d is not supported (the lunar element beam needs "
                    "CSPICE)\n", b->doBeam);
    exit(1);
  }
  (void)freqs_host;
  auto keep = [&](void *p) { bd->owned.push_back(p); return p; };
  BeamArgs g;
  memset(&g, 0, sizeof(g));
  g.src = sky.src; g.S = sky.nsrc; g.freqs = freqs_dev; g.Nf = Nf; g.f0 = b->ph_freq0;
  g.T = b->tilesz; g.N = N; g.bf_type = b->bf_type; g.b_ra0 = b->b_ra0; g.b_dec0 = b->b_dec0;
  g.ra0 = b->ph_ra0; g.dec0 = b->ph_dec0; g.wideband = wide ? 1 : 0;
  g.time_jd = (double *)keep(upload_doubles(b->time_utc, b->tilesz, st));
  g.lon = (double *)keep(upload_doubles(b->longitude, N, st));
  g.lat = (double *)keep(upload_doubles(b->latitude, N, st));
  const size_t ntab = (size_t)b->tilesz * Nf * sky.nsrc * N;
  if (do_array) {
    if (b->bf_type != STAT_SINGLE && b->bf_type != STAT_TILE) {
      fprintf(stderr, "dirac_b200: array beam needs bf_type STAT_SINGLE or STAT_TILE\n");
      exit(1);
    }
    std::vector<int> off(N), ne(N);
    std::vector<double> ex, ey, ez;
    for (int n = 0; n < N; n++) {
      off[n] = (int)ex.size();
      ne[n] = b->Nelem[n];
      const int len = b->Nelem[n] + (b->bf_type == STAT_TILE ? HBA_TILE_SIZE : 0);
      ex.insert(ex.end(), b->xx[n], b->xx[n] + len);
      ey.insert(ey.end(), b->yy[n], b->yy[n] + len);
      ez.insert(ez.end(), b->zz[n], b->zz[n] + len);
    }
    int *doff = nullptr, *dne = nullptr;
    DB_CHECK(cudaMalloc((void **)&doff, sizeof(int) * N));
    DB_CHECK(cudaMalloc((void **)&dne, sizeof(int) * N));
    DB_CHECK(cudaMemcpyAsync(doff, off.data(), sizeof(int) * N, cudaMemcpyHostToDevice, st));
    DB_CHECK(cudaMemcpyAsync(dne, ne.data(), sizeof(int) * N, cudaMemcpyHostToDevice, st));
    keep(doff); keep(dne);
    g.elem_off = doff; g.Nelem = dne;
    g.ex = (double *)keep(upload_doubles(ex.data(), (long long)ex.size(), st));
    g.ey = (double *)keep(upload_doubles(ey.data(), (long long)ey.size(), st));
    g.ez = (double *)keep(upload_doubles(ez.data(), (long long)ez.size(), st));
    DB_CHECK(cudaMalloc((void **)&bd->af, sizeof(double) * ntab));
    g.af = bd->af;
    db_stream_sync(st);  // the host vectors go out of scope
  }
  if (do_elem) {
    const elementcoeff *ec = b->ecoeff;
    if (!ec || !ec->pattern_phi || !ec->pattern_theta || !ec->preamble) {
      fprintf(stderr, "dirac_b200: element beam requested (doBeam %d) without coefficient tables "
                      "(set_elementcoeffs)\n", b->doBeam);
      exit(1);
    }
    const int nfc = wide ? ec->Nf : 1;
    g.ecM = ec->M; g.ecNmodes = ec->Nmodes; g.ecbeta = ec->beta;
    g.pat_phi = (double2 *)keep(upload_doubles(ec->pattern_phi, 2ll * ec->Nmodes * nfc, st));
    g.pat_theta = (double2 *)keep(upload_doubles(ec->pattern_theta, 2ll * ec->Nmodes * nfc, st));
    g.preamble = (double *)keep(upload_doubles(ec->preamble, ec->Nmodes, st));
    DB_CHECK(cudaMalloc((void **)&bd->E, sizeof(double2) * 4 * ntab));
    g.E = bd->E;
  }
  db_launch_beam_tables(&g, st);
  db_count_launch(1);
  // the coherency kernel needs the stations of every row now
  if (!a->sta1) {
    std::vector<int> s1(R), s2(R);
    for (long long r = 0; r < R; r++) {
      s1[r] = barr[r].sta1;
      s2[r] = barr[r].sta2;
    }
    int *ds1 = nullptr, *ds2 = nullptr;
    DB_CHECK(cudaMalloc((void **)&ds1, sizeof(int) * R));
    DB_CHECK(cudaMalloc((void **)&ds2, sizeof(int) * R));
    DB_CHECK(cudaMemcpyAsync(ds1, s1.data(), sizeof(int) * R, cudaMemcpyHostToDevice, st));
    DB_CHECK(cudaMemcpyAsync(ds2, s2.data(), sizeof(int) * R, cudaMemcpyHostToDevice, st));
    db_stream_sync(st);
    keep(ds1); keep(ds2);
    a->sta1 = ds1; a->sta2 = ds2;
  }
  a->beam_af = bd->af; a->beam_E = bd->E; a->beam_S = sky.nsrc; a->Nbase = Nbase_slot; a->N = N;
}
static void beam_free(BeamDev *bd) {
  if (bd->af) cudaFree(bd->af);
  if (bd->E) cudaFree(bd->E);
  for (void *p : bd->owned) cudaFree(p);
  bd->owned.clear();
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
  a.u = du; a.v = dv; a.w = dw; a.src = sky.src; a.modes = sky.modes; a.segs = sky.segs; a.nseg = sky.nseg;
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
static int precalculate_impl(double *u, double *v, double *w, double *x, int N, int Nbase,
                            baseline_t *barr, clus_source_t *carr, int M, double freq0,
                            double fdelta, double uvmin, double uvmax, const BeamSpec *beam) {
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
  a.u = du; a.v = dv; a.w = dw; a.src = sky.src; a.modes = sky.modes; a.segs = sky.segs; a.nseg = sky.nseg;
  a.freqs = df; a.Nchan = 1; a.fdelta2 = fdelta * 0.5; a.uvmin = uvmin; a.uvmax = uvmax;
  a.R = R; a.coh = dcoh; a.flag = dflag; a.xout = nullptr;
  BeamDev bd;
  beam_prepare(beam, sky, N, N * (N - 1) / 2, &freq0, 1, df, barr, R, st, &a, &bd);
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
  beam_free(&bd);
  sky_free(&sky);
  cudaStreamDestroy(st);
  return 0;
}
extern "C" int precalculate_coherencies(double *u, double *v, double *w, double *x, int N,
                                        int Nbase, baseline_t *barr, clus_source_t *carr, int M,
                                        double freq0, double fdelta, double tdelta, double dec0,
                                        double uvmin, double uvmax, int Nt) {
  (void)tdelta; (void)dec0; (void)Nt;
  return precalculate_impl(u, v, w, x, N, Nbase, barr, carr, M, freq0, fdelta, uvmin, uvmax, nullptr);
}
// Dirac_radio.h:472,516 (predict_withbeam.c:553-723): the same with the station beam towards every
// source folded in: array factor (a real gain per station) and / or element beam (a 2x2 E-Jones per
// station), evaluated per timeslot.  Nbase is Nbase*tilesz here too; rows in time order.
extern "C" int precalculate_coherencies_withbeam(
    double *u, double *v, double *w, double *x, int N, int Nbase, baseline_t *barr,
    clus_source_t *carr, int M, double freq0, double fdelta, double tdelta, double dec0, double uvmin,
    double uvmax, int bf_type, double b_ra0, double b_dec0, double ph_ra0, double ph_dec0,
    double ph_freq0, double *longitude, double *latitude, double *time_utc, int tilesz, int *Nelem,
    double **xx, double **yy, double **zz, elementcoeff *ecoeff, int doBeam, int Nt) {
  (void)tdelta; (void)dec0; (void)Nt;
  BeamSpec b = {bf_type, b_ra0, b_dec0, ph_ra0, ph_dec0, ph_freq0, longitude, latitude, time_utc,
                tilesz, Nelem, xx, yy, zz, ecoeff, doBeam};
  return precalculate_impl(u, v, w, x, N, Nbase, barr, carr, M, freq0, fdelta, uvmin, uvmax, &b);
}
extern "C" int precalculate_coherencies_withbeam_gpu(
    double *u, double *v, double *w, double *x, int N, int Nbase, baseline_t *barr,
    clus_source_t *carr, int M, double freq0, double fdelta, double tdelta, double dec0, double uvmin,
    double uvmax, int bf_type, double b_ra0, double b_dec0, double ph_ra0, double ph_dec0,
    double ph_freq0, double *longitude, double *latitude, double *time_utc, int tilesz, int *Nelem,
    double **xx, double **yy, double **zz, elementcoeff *ecoeff, int doBeam, int Nt) {
  return precalculate_coherencies_withbeam(u, v, w, x, N, Nbase, barr, carr, M, freq0, fdelta, tdelta,
                                           dec0, uvmin, uvmax, bf_type, b_ra0, b_dec0, ph_ra0,
                                           ph_dec0, ph_freq0, longitude, latitude, time_utc, tilesz,
                                           Nelem, xx, yy, zz, ecoeff, doBeam, Nt);
}

// Dirac_radio.h:221,479,534 (predict.c:745-816, predict_withbeam.c:726-900): coherencies of Nchan
// channels, x[chan][row][cluster][4] -- what the minibatch drivers feed bfgsfit_minibatch_*.  Fluxes
// are the catalogue's (no spectral index here, predict.c:690-696), the smearing width is
// fdelta / Nchan per channel (:792), a row gets flag 2 if it is shorter than uvmin at the first
// channel or longer than uvmax at the last (:731-735); the beam variant cuts both ways at the
// beam-former's reference frequency instead (predict_withbeam.c:456-463,784).  One pass of the single-channel kernel per
// channel; wide-band element beams see their channel's coefficient set.
static int precalculate_multifreq_impl(double *u, double *v, double *w, double *x, int N, int Nbase,
                                       baseline_t *barr, clus_source_t *carr, int M, double *freqs,
                                       int Nchan, double fdelta, double uvmin, double uvmax,
                                       const BeamSpec *beam) {
  const double HUGE_UV = 1e300;
  for (int c = 0; c < Nchan; c++) {
    // (the beam variant cuts on ONE frequency, ph_freq0: done on the host below)
    const double lo = (c == 0 && !beam) ? uvmin : 0.0;
    const double hi = (c == Nchan - 1 && !beam) ? uvmax : HUGE_UV;
    BeamSpec bc;
    elementcoeff ec;
    const BeamSpec *bp = nullptr;
    if (beam) {
      bc = *beam;
      const bool wide = beam->doBeam == DOBEAM_ARRAY_WB || beam->doBeam == DOBEAM_FULL_WB ||
                        beam->doBeam == DOBEAM_ELEMENT_WB;
      if (wide && beam->ecoeff) {  // this channel's coefficient set as a one-frequency table
        ec = *beam->ecoeff;
        ec.pattern_phi = beam->ecoeff->pattern_phi + (size_t)2 * ec.Nmodes * c;
        ec.pattern_theta = beam->ecoeff->pattern_theta + (size_t)2 * ec.Nmodes * c;
        ec.Nf = 1;
        bc.ecoeff = &ec;
      }
      bp = &bc;
    }
    const int rv = precalculate_impl(u, v, w, x + (size_t)c * 8 * M * Nbase, N, Nbase, barr, carr, M,
                                     freqs[c], fdelta / (double)Nchan, lo, hi, bp);
    if (rv) return rv;
  }
  if (beam)  // predict_withbeam.c:456-463 with freq0 = ph_freq0 (:784)
    for (long long r = 0; r < Nbase; r++)
      if (!barr[r].flag) {
        const double uvdist = sqrt(u[r] * u[r] + v[r] * v[r]) * beam->ph_freq0;
        if (uvdist < uvmin || uvdist > uvmax) barr[r].flag = 2;
      }
  return 0;
}
extern "C" int precalculate_coherencies_multifreq(double *u, double *v, double *w, double *x, int N,
                                                  int Nbase, baseline_t *barr, clus_source_t *carr,
                                                  int M, double *freqs, int Nchan, double fdelta,
                                                  double tdelta, double dec0, double uvmin,
                                                  double uvmax, int Nt) {
  (void)tdelta; (void)dec0; (void)Nt;
  return precalculate_multifreq_impl(u, v, w, x, N, Nbase, barr, carr, M, freqs, Nchan, fdelta, uvmin,
                                     uvmax, nullptr);
}
extern "C" int precalculate_coherencies_multifreq_withbeam(
    double *u, double *v, double *w, double *x, int N, int Nbase, baseline_t *barr,
    clus_source_t *carr, int M, double *freqs, int Nchan, double fdelta, double tdelta, double dec0,
    double uvmin, double uvmax, int bf_type, double b_ra0, double b_dec0, double ph_ra0,
    double ph_dec0, double ph_freq0, double *longitude, double *latitude, double *time_utc, int tilesz,
    int *Nelem, double **xx, double **yy, double **zz, elementcoeff *ecoeff, int doBeam, int Nt) {
  (void)tdelta; (void)dec0; (void)Nt;
  BeamSpec b = {bf_type, b_ra0, b_dec0, ph_ra0, ph_dec0, ph_freq0, longitude, latitude, time_utc,
                tilesz, Nelem, xx, yy, zz, ecoeff, doBeam};
  return precalculate_multifreq_impl(u, v, w, x, N, Nbase, barr, carr, M, freqs, Nchan, fdelta, uvmin,
                                     uvmax, &b);
}
extern "C" int precalculate_coherencies_multifreq_withbeam_gpu(
    double *u, double *v, double *w, double *x, int N, int Nbase, baseline_t *barr,
    clus_source_t *carr, int M, double *freqs, int Nchan, double fdelta, double tdelta, double dec0,
    double uvmin, double uvmax, int bf_type, double b_ra0, double b_dec0, double ph_ra0,
    double ph_dec0, double ph_freq0, double *longitude, double *latitude, double *time_utc, int tilesz,
    int *Nelem, double **xx, double **yy, double **zz, elementcoeff *ecoeff, int doBeam, int Nt) {
  return precalculate_coherencies_multifreq_withbeam(u, v, w, x, N, Nbase, barr, carr, M, freqs, Nchan,
                                                     fdelta, tdelta, dec0, uvmin, uvmax, bf_type,
                                                     b_ra0, b_dec0, ph_ra0, ph_dec0, ph_freq0,
                                                     longitude, latitude, time_utc, tilesz, Nelem, xx,
                                                     yy, zz, ecoeff, doBeam, Nt);
}

// Dirac_radio.h:659 (residual.c:1257-1340): x[chan][row][8] += sum over clusters; add_to_data ==
// SIMUL_ONLY (1, Dirac_radio.h:78) clears x first, every other value accumulates onto the input
// (the thread function only ever adds, residual.c:1238-1245).  No Jones, no flags.
static int predict_multifreq_impl(double *u, double *v, double *w, double *x, int N, int Nbase,
                                  int tilesz, baseline_t *barr, clus_source_t *carr, int M,
                                  double *freqs, int Nchan, double fdelta, int add_to_data,
                                  const BeamSpec *beam) {
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
  a.u = du; a.v = dv; a.w = dw; a.src = sky.src; a.modes = sky.modes; a.segs = sky.segs; a.nseg = sky.nseg;
  a.freqs = df; a.Nchan = Nchan; a.fdelta2 = (fdelta / (double)Nchan) * 0.5;
  a.R = R; a.xout = dx;
  BeamDev bd;
  beam_prepare(beam, sky, N, Nbase, freqs, Nchan, df, barr, R, st, &a, &bd);
  db_launch_predict_multifreq(&a, st);
  db_count_launch(1);
  DB_CHECK(cudaMemcpyAsync(x, dx, sizeof(double2) * nx, cudaMemcpyDeviceToHost, st));
  db_stream_sync(st);
  DB_CHECK(cudaGetLastError());
  cudaFree(dx); cudaFree(du); cudaFree(dv); cudaFree(dw); cudaFree(df);
  beam_free(&bd);
  sky_free(&sky);
  cudaStreamDestroy(st);
  return 0;
}
extern "C" int predict_visibilities_multifreq(double *u, double *v, double *w, double *x, int N,
                                              int Nbase, int tilesz, baseline_t *barr,
                                              clus_source_t *carr, int M, double *freqs, int Nchan,
                                              double fdelta, double tdelta, double dec0, int Nt,
                                              int add_to_data) {
  (void)tdelta; (void)dec0; (void)Nt;
  return predict_multifreq_impl(u, v, w, x, N, Nbase, tilesz, barr, carr, M, freqs, Nchan, fdelta,
                                add_to_data, nullptr);
}
// Dirac_radio.h:485,521 (predict_withbeam.c:1219-1440): per-channel station beams
extern "C" int predict_visibilities_multifreq_withbeam(
    double *u, double *v, double *w, double *x, int N, int Nbase, int tilesz, baseline_t *barr,
    clus_source_t *carr, int M, double *freqs, int Nchan, double fdelta, double tdelta, double dec0,
    int bf_type, double b_ra0, double b_dec0, double ph_ra0, double ph_dec0, double ph_freq0,
    double *longitude, double *latitude, double *time_utc, int *Nelem, double **xx, double **yy,
    double **zz, elementcoeff *ecoeff, int doBeam, int Nt, int add_to_data) {
  (void)tdelta; (void)dec0; (void)Nt;
  BeamSpec b = {bf_type, b_ra0, b_dec0, ph_ra0, ph_dec0, ph_freq0, longitude, latitude, time_utc,
                tilesz, Nelem, xx, yy, zz, ecoeff, doBeam};
  return predict_multifreq_impl(u, v, w, x, N, Nbase, tilesz, barr, carr, M, freqs, Nchan, fdelta,
                                add_to_data, &b);
}
extern "C" int predict_visibilities_multifreq_withbeam_gpu(
    double *u, double *v, double *w, double *x, int N, int Nbase, int tilesz, baseline_t *barr,
    clus_source_t *carr, int M, double *freqs, int Nchan, double fdelta, double tdelta, double dec0,
    int bf_type, double b_ra0, double b_dec0, double ph_ra0, double ph_dec0, double ph_freq0,
    double *longitude, double *latitude, double *time_utc, int *Nelem, double **xx, double **yy,
    double **zz, elementcoeff *ecoeff, int doBeam, int Nt, int add_to_data) {
  return predict_visibilities_multifreq_withbeam(u, v, w, x, N, Nbase, tilesz, barr, carr, M, freqs,
                                                 Nchan, fdelta, tdelta, dec0, bf_type, b_ra0, b_dec0,
                                                 ph_ra0, ph_dec0, ph_freq0, longitude, latitude,
                                                 time_utc, Nelem, xx, yy, zz, ecoeff, doBeam, Nt,
                                                 add_to_data);
}

// 2x2 inverse of (J + rho I) with the reference's guard on a small determinant (mat_invert,
// residual.c:162-199)
static void jones_invert(const double xx[8], double yy[8], double rho) {
  const double a0r = xx[0] + rho, a0i = xx[1], a1r = xx[2], a1i = xx[3];
  const double a2r = xx[4], a2i = xx[5], a3r = xx[6] + rho, a3i = xx[7];
  double dr = (a0r * a3r - a0i * a3i) - (a1r * a2r - a1i * a2i);
  double di = (a0r * a3i + a0i * a3r) - (a1r * a2i + a1i * a2r);
  if (sqrt(sqrt(dr * dr + di * di)) <= rho) dr += rho;
  const double den = dr * dr + di * di;
  const double ir = dr / den, ii = -di / den;  // 1/det
  yy[0] = a3r * ir - a3i * ii;      yy[1] = a3r * ii + a3i * ir;
  yy[2] = -(a1r * ir - a1i * ii);   yy[3] = -(a1r * ii + a1i * ir);
  yy[4] = -(a2r * ir - a2i * ii);   yy[5] = -(a2r * ii + a2i * ir);
  yy[6] = a0r * ir - a0i * ii;      yy[7] = a0r * ii + a0i * ir;
}

// eigenvector of the largest eigenvalue of a real symmetric 3x3 matrix by cyclic Jacobi rotations
// (stands in for dsyevx with IL = IU = 3, manifold_average.c:470,541)
static void sym3_top_eigvec(const double Hin[3][3], double z[3]) {
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A[i][j] = Hin[i][j];
  for (int sweep = 0; sweep < 60; sweep++) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double dg = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-34 * dg || off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 3; k++) {  // A <- A R
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - sn * akq;
          A[k][q] = sn * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {  // A <- R^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - sn * aqk;
          A[q][k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - sn * vkq;
          V[k][q] = sn * vkp + c * vkq;
        }
      }
  }
  int top = 0;
  for (int i = 1; i < 3; i++)
    if (A[i][i] > A[top][top]) top = i;
  const double nrm = sqrt(V[0][top] * V[0][top] + V[1][top] * V[1][top] + V[2][top] * V[2][top]);
  for (int i = 0; i < 3; i++) z[i] = V[i][top] / nrm;
}

// phases of the jointly diagonalised solutions of one (cluster, chunk): niter rounds of two Jacobi
// rotations J <- J G^H common to all stations (towards diagonal J), then the unit-modulus diagonal
// (extract_phases, manifold_average.c:399-610).  pin / pout: N x 8 doubles in the pp layout.
static void extract_phases_host(const double *pin, double *pout, int N, int niter) {
  struct cd { double r, i; };
  auto mul = [](cd a, cd b) { return cd{a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; };
  auto conj = [](cd a) { return cd{a.r, -a.i}; };
  std::vector<cd> J00(N), J01(N), J10(N), J11(N);
  for (int s = 0; s < N; s++) {
    J00[s] = {pin[8 * s + 0], pin[8 * s + 1]};
    J01[s] = {pin[8 * s + 2], pin[8 * s + 3]};
    J10[s] = {pin[8 * s + 4], pin[8 * s + 5]};
    J11[s] = {pin[8 * s + 6], pin[8 * s + 7]};
  }
  for (int ni = 0; ni < niter; ni++)
    for (int pass = 0; pass < 2; pass++) {
      double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (int s = 0; s < N; s++) {
        // pass 0: h = conj[a - d, b + c, i (c - b)]; pass 1: h = conj[d - a, c + b, i (b - c)]
        const cd a = J00[s], b = J01[s], c = J10[s], d = J11[s];
        const double sg = pass == 0 ? 1.0 : -1.0;
        cd h[3];
        h[0] = conj(cd{sg * (a.r - d.r), sg * (a.i - d.i)});
        h[1] = conj(cd{b.r + c.r, b.i + c.i});
        const cd cmb = {sg * (c.r - b.r), sg * (c.i - b.i)};
        h[2] = conj(cd{-cmb.i, cmb.r});  // i (c - b)
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) H[i][j] += h[i].r * h[j].r + h[i].i * h[j].i;  // Re(h h^H)
      }
      double Z[3];
      sym3_top_eigvec(H, Z);
      cd cc, ss;
      if (Z[0] >= 0.0) {
        cc = {sqrt(0.5 + Z[0] * 0.5), 0.0};
        ss = {0.5 * Z[1] / cc.r, -0.5 * Z[2] / cc.r};
      } else {
        cc = {sqrt(0.5 - Z[0] * 0.5), 0.0};
        ss = {-0.5 * Z[1] / cc.r, 0.5 * Z[2] / cc.r};
      }
      // G = [c, conj(s); -s, conj(c)] (column-major [c, -s, conj(s), conj(c)]);  J <- J G^H
      // (G^H)(0,0) = conj(c), (G^H)(0,1) = conj(-s), (G^H)(1,0) = s, (G^H)(1,1) = c
      const cd g00 = conj(cc), g01 = conj(cd{-ss.r, -ss.i}), g10 = ss, g11 = cc;
      for (int s = 0; s < N; s++) {
        const cd a = J00[s], b = J01[s], c = J10[s], d = J11[s];
        const cd t0 = mul(a, g00), t1 = mul(b, g10), t2 = mul(a, g01), t3 = mul(b, g11);
        const cd u0 = mul(c, g00), u1 = mul(d, g10), u2 = mul(c, g01), u3 = mul(d, g11);
        J00[s] = {t0.r + t1.r, t0.i + t1.i};
        J01[s] = {t2.r + t3.r, t2.i + t3.i};
        J10[s] = {u0.r + u1.r, u0.i + u1.i};
        J11[s] = {u2.r + u3.r, u2.i + u3.i};
      }
    }
  memset(pout, 0, sizeof(double) * 8 * N);
  for (int s = 0; s < N; s++) {
    const double m0 = sqrt(J00[s].r * J00[s].r + J00[s].i * J00[s].i);
    const double m1 = sqrt(J11[s].r * J11[s].r + J11[s].i * J11[s].i);
    pout[8 * s + 0] = J00[s].r / m0;
    pout[8 * s + 1] = J00[s].i / m0;
    pout[8 * s + 6] = J11[s].r / m1;
    pout[8 * s + 7] = J11[s].i / m1;
  }
}

// host arithmetic, no GPU needed: exposed so that the CPU tests can pin it against the reference
extern "C" int dirac_b200_extract_phases(const double *p, double *pout, int N, int niter) {
  extract_phases_host(p, pout, N, niter);
  return 0;
}

// Dirac_radio.h:652 (residual.c:940-1061): full-resolution residual per channel,
// x[chan][row][8] -= sum over the clusters with id >= 0 of J_p C_k(chan) J_q^H, the coherencies
// re-predicted from the sources at every channel frequency; then, if a cluster has id == ccid, every
// row is corrected by that cluster's inverse Jones (J + rho I)^-1; phase_only != 0: by the inverse of
// the phases of its jointly diagonalised solutions (extract_phases, manifold_average.c:399-610).
static int residuals_multifreq_impl(double *u, double *v, double *w, double *p, double *x, int N,
                                    int Nbase, int tilesz, baseline_t *barr, clus_source_t *carr,
                                    int M, double *freqs, int Nchan, double fdelta, int ccid,
                                    double rho, int phase_only, const BeamSpec *beam) {
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
  // cluster / chunk tables, stations, solutions
  std::vector<int> nchunk(M), chunk0(M), poff, s1(R), s2(R);
  std::vector<unsigned char> sub(M);
  int mt = 0, cm = -1;
  long long npar = 0;
  for (int k = 0; k < M; k++) {
    nchunk[k] = carr[k].nchunk;
    chunk0[k] = mt;
    sub[k] = carr[k].id >= 0 ? 1 : 0;
    if (carr[k].id == ccid) cm = k;
    for (int c = 0; c < carr[k].nchunk; c++) {
      poff.push_back(carr[k].p[c]);
      if ((long long)carr[k].p[c] + 8ll * N > npar) npar = (long long)carr[k].p[c] + 8ll * N;
    }
    mt += carr[k].nchunk;
  }
  for (long long r = 0; r < R; r++) {
    s1[r] = barr[r].sta1;
    s2[r] = barr[r].sta2;
  }
  std::vector<double> pinv;
  if (cm >= 0) {
    pinv.resize((size_t)8 * N * carr[cm].nchunk);
    std::vector<double> pphase(phase_only ? (size_t)8 * N : 0);
    for (int c = 0; c < carr[cm].nchunk; c++) {
      const double *pm = p + carr[cm].p[c];
      if (phase_only) {  // only the phases of the jointly diagonalised solutions (residual.c:975-990)
        extract_phases_host(pm, pphase.data(), N, 10);
        pm = pphase.data();
      }
      for (int s = 0; s < N; s++)
        jones_invert(pm + 8 * s, pinv.data() + (size_t)8 * N * c + 8 * s, rho);
    }
  }
  int *dn = nullptr, *dc0 = nullptr, *dpo = nullptr, *ds1 = nullptr, *ds2 = nullptr;
  unsigned char *dsub = nullptr;
  DB_CHECK(cudaMalloc((void **)&dn, sizeof(int) * M));
  DB_CHECK(cudaMalloc((void **)&dc0, sizeof(int) * M));
  DB_CHECK(cudaMalloc((void **)&dpo, sizeof(int) * (mt > 0 ? mt : 1)));
  DB_CHECK(cudaMalloc((void **)&ds1, sizeof(int) * R));
  DB_CHECK(cudaMalloc((void **)&ds2, sizeof(int) * R));
  DB_CHECK(cudaMalloc((void **)&dsub, M));
  DB_CHECK(cudaMemcpyAsync(dn, nchunk.data(), sizeof(int) * M, cudaMemcpyHostToDevice, st));
  DB_CHECK(cudaMemcpyAsync(dc0, chunk0.data(), sizeof(int) * M, cudaMemcpyHostToDevice, st));
  DB_CHECK(cudaMemcpyAsync(dpo, poff.data(), sizeof(int) * mt, cudaMemcpyHostToDevice, st));
  DB_CHECK(cudaMemcpyAsync(ds1, s1.data(), sizeof(int) * R, cudaMemcpyHostToDevice, st));
  DB_CHECK(cudaMemcpyAsync(ds2, s2.data(), sizeof(int) * R, cudaMemcpyHostToDevice, st));
  DB_CHECK(cudaMemcpyAsync(dsub, sub.data(), M, cudaMemcpyHostToDevice, st));
  double *dp = upload_doubles(p, npar, st);
  double *dpinv = cm >= 0 ? upload_doubles(pinv.data(), (long long)pinv.size(), st) : nullptr;
  double2 *dx = nullptr;
  const size_t nx = (size_t)Nchan * R * 4;
  DB_CHECK(cudaMalloc((void **)&dx, sizeof(double2) * nx));
  DB_CHECK(cudaMemcpyAsync(dx, x, sizeof(double2) * nx, cudaMemcpyHostToDevice, st));
  CohArgs a;
  memset(&a, 0, sizeof(a));
  a.u = du; a.v = dv; a.w = dw; a.src = sky.src; a.modes = sky.modes; a.segs = sky.segs; a.nseg = sky.nseg;
  a.freqs = df; a.Nchan = Nchan; a.fdelta2 = (fdelta / (double)Nchan) * 0.5;
  a.R = R; a.xout = dx; a.sta1 = ds1; a.sta2 = ds2; a.p = dp; a.clus_nchunk = dn;
  a.clus_chunk0 = dc0; a.chunk_poff = dpo; a.clus_sub = dsub; a.pinv = dpinv;
  a.pinv_nchunk = cm >= 0 ? carr[cm].nchunk : 1; a.N = N;
  BeamDev bd;
  beam_prepare(beam, sky, N, Nbase, freqs, Nchan, df, barr, R, st, &a, &bd);
  db_launch_residual_multifreq(&a, st);
  db_count_launch(1);
  DB_CHECK(cudaMemcpyAsync(x, dx, sizeof(double2) * nx, cudaMemcpyDeviceToHost, st));
  db_stream_sync(st);
  DB_CHECK(cudaGetLastError());
  cudaFree(dx); cudaFree(du); cudaFree(dv); cudaFree(dw); cudaFree(df); cudaFree(dp);
  if (dpinv) cudaFree(dpinv);
  cudaFree(dn); cudaFree(dc0); cudaFree(dpo); cudaFree(ds1); cudaFree(ds2); cudaFree(dsub);
  beam_free(&bd);
  sky_free(&sky);
  cudaStreamDestroy(st);
  return 0;
}
extern "C" int calculate_residuals_multifreq(double *u, double *v, double *w, double *p, double *x,
                                             int N, int Nbase, int tilesz, baseline_t *barr,
                                             clus_source_t *carr, int M, double *freqs, int Nchan,
                                             double fdelta, double tdelta, double dec0, int Nt,
                                             int ccid, double rho, int phase_only) {
  (void)tdelta; (void)dec0; (void)Nt;
  return residuals_multifreq_impl(u, v, w, p, x, N, Nbase, tilesz, barr, carr, M, freqs, Nchan, fdelta,
                                  ccid, rho, phase_only, nullptr);
}
// Dirac_radio.h:489,525 (predict_withbeam.c:1989-2315): per-channel station beams in the re-prediction
extern "C" int calculate_residuals_multifreq_withbeam(
    double *u, double *v, double *w, double *p, double *x, int N, int Nbase, int tilesz,
    baseline_t *barr, clus_source_t *carr, int M, double *freqs, int Nchan, double fdelta,
    double tdelta, double dec0, int bf_type, double b_ra0, double b_dec0, double ph_ra0,
    double ph_dec0, double ph_freq0, double *longitude, double *latitude, double *time_utc,
    int *Nelem, double **xx, double **yy, double **zz, elementcoeff *ecoeff, int doBeam, int Nt,
    int ccid, double rho, int phase_only) {
  (void)tdelta; (void)dec0; (void)Nt;
  BeamSpec b = {bf_type, b_ra0, b_dec0, ph_ra0, ph_dec0, ph_freq0, longitude, latitude, time_utc,
                tilesz, Nelem, xx, yy, zz, ecoeff, doBeam};
  return residuals_multifreq_impl(u, v, w, p, x, N, Nbase, tilesz, barr, carr, M, freqs, Nchan, fdelta,
                                  ccid, rho, phase_only, &b);
}
extern "C" int calculate_residuals_multifreq_withbeam_gpu(
    double *u, double *v, double *w, double *p, double *x, int N, int Nbase, int tilesz,
    baseline_t *barr, clus_source_t *carr, int M, double *freqs, int Nchan, double fdelta,
    double tdelta, double dec0, int bf_type, double b_ra0, double b_dec0, double ph_ra0,
    double ph_dec0, double ph_freq0, double *longitude, double *latitude, double *time_utc,
    int *Nelem, double **xx, double **yy, double **zz, elementcoeff *ecoeff, int doBeam, int Nt,
    int ccid, double rho, int phase_only) {
  return calculate_residuals_multifreq_withbeam(u, v, w, p, x, N, Nbase, tilesz, barr, carr, M, freqs,
                                                Nchan, fdelta, tdelta, dec0, bf_type, b_ra0, b_dec0,
                                                ph_ra0, ph_dec0, ph_freq0, longitude, latitude,
                                                time_utc, Nelem, xx, yy, zz, ecoeff, doBeam, Nt, ccid,
                                                rho, phase_only);
}
