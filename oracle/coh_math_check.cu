// TEST INFRASTRUCTURE ONLY (oracle).  Runs the per-source / per-station arithmetic of the product's
// coherency kernels (sagecal_b200/csrc/coh_math.cuh: shapelet factor, array factor, element beam) on
// the CPU so that tests/test_oracle_coh_math.py can pin it against the compiled reference's
// shapelet_contrib, arraybeam, array_element_beam and element_beam without a GPU.  Compiled by nvcc as
// host code; nothing is launched.
#include <string.h>
#include <vector>

#include "../sagecal_b200/csrc/coh_math.cuh"

// shapelet factor of one source at (u, v, w) already multiplied by the frequency (shapelet.c:141-190)
extern "C" void check_shapelet(int n0, double beta, const double *modes, double eX, double eY,
                               double eP, double cxi, double sxi, double cphi, double sphi,
                               int use_projection, double uf, double vf, double wf, double *out2) {
  DevSource s;
  memset(&s, 0, sizeof(s));
  s.eX = eX; s.eY = eY; s.eP = eP; s.cxi = cxi; s.sxi = sxi; s.cphi = cphi; s.sphi = sphi;
  s.use_projection = (double)use_projection;
  s.sh_n0 = (double)n0; s.sh_beta = beta; s.sh_off = 0.0;
  const double2 v = shapelet_factor(s, modes, uf, vf, wf);
  out2[0] = v.x;
  out2[1] = v.y;
}

// beam tables towards ONE source at ONE time and frequency for all N stations: what k_beam_tables
// computes per thread (stationbeam.c:49-430).  elem: packed element positions, off[n] first element of
// station n.  af[N] and / or E[N][8] out (null: skipped).
extern "C" void check_beam(double ra, double dec, int bf_type, double b_ra0, double b_dec0,
                           double ra0, double dec0, double f, double f0, int N, const double *lon,
                           const double *lat, double time_jd, const int *Nelem, const int *off,
                           const double *ex, const double *ey, const double *ez, int wideband,
                           int ecM, int ecNmodes, double ecbeta, const double *pat_phi,
                           const double *pat_theta, const double *preamble, int findex, double *af,
                           double *E) {
  DevSource s;
  memset(&s, 0, sizeof(s));
  s.ra = ra;
  s.dec = dec;
  // one "channel" whose coefficient set is findex: shift the tables instead of the index
  BeamArgs a;
  memset(&a, 0, sizeof(a));
  a.src = &s; a.S = 1; a.freqs = &f; a.Nf = 1; a.f0 = f0; a.time_jd = &time_jd; a.T = 1;
  a.lon = lon; a.lat = lat; a.N = N; a.elem_off = off; a.Nelem = Nelem; a.ex = ex; a.ey = ey; a.ez = ez;
  a.bf_type = bf_type; a.b_ra0 = b_ra0; a.b_dec0 = b_dec0; a.ra0 = ra0; a.dec0 = dec0;
  a.wideband = wideband;
  a.ecM = ecM; a.ecNmodes = ecNmodes; a.ecbeta = ecbeta;
  a.pat_phi = pat_phi ? reinterpret_cast<const double2 *>(pat_phi) + (size_t)findex * ecNmodes : nullptr;
  a.pat_theta = pat_theta ? reinterpret_cast<const double2 *>(pat_theta) + (size_t)findex * ecNmodes : nullptr;
  a.preamble = preamble;
  a.af = af;
  a.E = reinterpret_cast<double2 *>(E);
  for (int n = 0; n < N; n++) beam_table_entry(a, (size_t)n);
}
