// Coherency prediction from the sky model on the device.
//
//   k_coherencies        one frequency, every cluster -> resident planar coh[k][4][R]
//                        (replaces precal_threadfn, predict.c:345-497)
//   k_predict_multifreq  sum over clusters per channel with spectral-index fluxes -> x[chan][row][8]
//                        (replaces visibilities_threadfn_multifreq, residual.c:1067-1248)
//
// One thread per row (u,v,w coalesced in, 16-byte planar stores out).  The sources of a direction
// are staged in shared memory by a 1-D TMA bulk copy (cp.async.bulk + mbarrier complete_tx),
// double buffered so the copy of direction k+1 overlaps the trigonometry of direction k.  This
// kernel is SFU/FP64-ALU bound (three sin/cos per source-row), not HBM bound.
#include "internal.cuh"
#include "coh.h"
#include "tma.cuh"

// ---- per-source term ------------------------------------------------------------------------------
// phase * |sinc| smearing * shape factor for one source at one frequency (predict.c:411-470)
// Fourier-plane value of a shapelet source (shapelet_contrib + calculate_uv_mode_vectors_scalar,
// shapelet.c:50-190): sum over the n0 x n0 modes of coeff * phi_n1(-ut beta) phi_n2(vt beta), odd
// n1+n2 imaginary, phi_n(x) = H_n(x) exp(-x^2/2) / sqrt(2^(n+1) n!), times 2 pi / (eX eY)
__device__ __noinline__ double2 shapelet_factor(const DevSource &s, const double *modes, double uf,
                                                double vf, double wf) {
  double up, vp;
  if (s.use_projection != 0.0) {
    up = -uf * s.cxi + vf * s.cphi * s.sxi - wf * s.sphi * s.sxi;
    vp = -uf * s.sxi - vf * s.cphi * s.cxi + wf * s.sphi * s.cxi;
  } else {
    up = uf;
    vp = vf;
  }
  const double a = 1.0 / s.eX, b = 1.0 / s.eY;
  double sph, cph;
  sincos(s.eP, &sph, &cph);
  const double ut = a * (cph * up - sph * vp);
  const double vt = b * (sph * up + cph * vp);
  const int n0 = (int)s.sh_n0;
  double bu[COH_SHAPELET_MAX_N0], bv[COH_SHAPELET_MAX_N0];
#pragma unroll 1
  for (int side = 0; side < 2; side++) {
    const double x = (side == 0 ? -ut : vt) * s.sh_beta;
    const double ex = exp(-0.5 * x * x);
    double *bb = side == 0 ? bu : bv;
    double hm2 = 1.0, hm1 = 2.0 * x, fact = 1.0, p2 = 2.0;  // H_0, H_1, n!, 2^(n+1)
    for (int n = 0; n < n0; n++) {
      double h;
      if (n == 0) h = 1.0;
      else if (n == 1) h = hm1;
      else {
        h = 2.0 * x * hm1 - 2.0 * (double)(n - 1) * hm2;
        hm2 = hm1;
        hm1 = h;
      }
      if (n > 0) fact *= (double)n;
      bb[n] = h * ex / sqrt(p2 * fact);
      p2 *= 2.0;
    }
  }
  const double *md = modes + (long long)s.sh_off;
  double re = 0.0, im = 0.0;
  for (int n2 = 0; n2 < n0; n2++)
    for (int n1 = 0; n1 < n0; n1++) {
      const int odd = (n1 + n2) & 1;
      const int sg = (((n1 + n2 - odd) / 2) & 1) ? -1 : 1;
      const double av = (sg < 0 ? -bu[n1] : bu[n1]) * bv[n2];
      const double c = md[n2 * n0 + n1] * av;
      if (odd) im += c;
      else re += c;
    }
  const double sc = 2.0 * M_PI * a * b;
  return make_double2(sc * re, sc * im);
}

__device__ __forceinline__ double2 source_phase(const DevSource &s, const double *modes, double u,
                                                double v, double w, double freq, double fdelta2) {
  const double G = 2.0 * M_PI * (u * s.ll + v * s.mm + w * s.nn);
  double sp, cp;
  sincos(G * freq, &sp, &cp);
  double fac = 1.0;
  if (G != 0.0) {
    const double sm = G * fdelta2;
    fac = fabs(sin(sm) / sm);
  }
  double2 ph = make_double2(cp * fac, sp * fac);
  const int st = (int)s.stype;
  if (st == STYPE_SHAPELET_) {
    const double2 sf = shapelet_factor(s, modes, u * freq, v * freq, w * freq);
    ph = make_double2(ph.x * sf.x - ph.y * sf.y, ph.x * sf.y + ph.y * sf.x);
  } else if (st != STYPE_POINT_) {
    const double uf = u * freq, vf = v * freq, wf = w * freq;
    double up, vp;
    if (st == STYPE_GAUSSIAN_ && s.use_projection == 0.0) {
      up = uf;
      vp = vf;
    } else {
      up = uf * s.cxi - vf * s.cphi * s.sxi + wf * s.sphi * s.sxi;
      vp = uf * s.sxi + vf * s.cphi * s.cxi - wf * s.sphi * s.cxi;
    }
    double shape = 1.0;
    if (st == STYPE_GAUSSIAN_) {
      double sph, cph;
      sincos(s.eP, &sph, &cph);
      const double ut = s.eX * (cph * up - sph * vp);
      const double vt = s.eY * (sph * up + cph * vp);
      shape = exp(-2.0 * M_PI * M_PI * (ut * ut + vt * vt));
    } else if (st == STYPE_DISK_) {
      shape = j1(sqrt(up * up + vp * vp) * s.eX * 2.0 * M_PI);
    } else if (st == STYPE_RING_) {
      shape = j0(sqrt(up * up + vp * vp) * s.eX * 2.0 * M_PI);
    }
    ph.x *= shape;
    ph.y *= shape;
  }
  return ph;
}

__device__ __forceinline__ void add_stokes(double2 *C, double2 ph, double I, double Q, double U,
                                           double V) {
  // C0 += ph (I+Q); C1 += ph (U + iV); C2 += ph (U - iV); C3 += ph (I-Q)   (predict.c:466-476)
  const double2 II = make_double2(ph.x * I, ph.y * I), QQ = make_double2(ph.x * Q, ph.y * Q);
  const double2 UU = make_double2(ph.x * U, ph.y * U), VV = make_double2(ph.x * V, ph.y * V);
  C[0].x += II.x + QQ.x;  C[0].y += II.y + QQ.y;
  C[1].x += UU.x - VV.y;  C[1].y += UU.y + VV.x;
  C[2].x += UU.x + VV.y;  C[2].y += UU.y - VV.x;
  C[3].x += II.x - QQ.x;  C[3].y += II.y - QQ.y;
}

// flux at frequency f with the three-term log-spectral index (residual.c:1177-1210)
__device__ __forceinline__ double spec_flux(double s0, double tempfr) {
  if (s0 > 0.0) return exp(log(s0) + tempfr);
  return (s0 == 0.0) ? 0.0 : -exp(log(-s0) + tempfr);
}

#define COH_THREADS 128

// MODE 0: coherencies per cluster at freq[0] -> planar coh ; MODE 1: multifreq sum -> xout ;
// MODE 2: xout -= sum_k J_p C_k J_q^H per channel with the solved Jones (clusters with id >= 0), then
//         the optional correction x <- Jinv_p x Jinv_q^H by one cluster's inverse Jones
//         (residual_threadfn_multifreq, residual.c:681-938)
template <int MODE>
__global__ void __launch_bounds__(COH_THREADS)
k_sky_predict(CohArgs a) {
  __shared__ __align__(128) DevSource sbuf[2][COH_SEG_MAX];
  __shared__ __align__(8) unsigned long long bar[2];
  const long long r = (long long)blockIdx.x * COH_THREADS + threadIdx.x;
  const bool active = r < a.R;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  double u = 0.0, v = 0.0, w = 0.0;
  if (active) {
    u = a.u[r];
    v = a.v[r];
    w = a.w[r];
  }
  const int nchan = (MODE == 0) ? 1 : a.Nchan;
  int s1 = 0, s2 = 0, tslot = 0;
  if ((MODE == 2 || a.sta1) && active) {
    s1 = a.sta1[r];
    s2 = a.sta2[r];
  }
  if (a.beam_af || a.beam_E) tslot = (int)(r / a.Nbase);
  // prologue: stage segment 0
  if (threadIdx.x == 0 && a.nseg > 0) {
    const CohSegment sg = a.segs[0];
    const unsigned bytes = (unsigned)sg.count * (unsigned)sizeof(DevSource);
    mbar_expect_tx(&bar[0], bytes);
    if (bytes) bulk_g2s(&sbuf[0][0], a.src + sg.first, bytes, &bar[0]);
  }
  for (int cf = 0; cf < nchan; cf++) {
    // (channels re-walk the segment list; the staging pipeline simply continues)
    double2 X[4];
#pragma unroll
    for (int c = 0; c < 4; c++) X[c] = make_double2(0.0, 0.0);
    const double freq = a.freqs[cf];
    double2 C[4];
#pragma unroll
    for (int c = 0; c < 4; c++) C[c] = make_double2(0.0, 0.0);
    for (int sgi = 0; sgi < a.nseg; sgi++) {
      const int it = cf * a.nseg + sgi;  // global staging iteration
      const int b = it & 1;
      // stage the next segment (of this or the next channel) into the other buffer
      if (threadIdx.x == 0) {
        int nxt = sgi + 1;
        bool more = true;
        if (nxt == a.nseg) {
          nxt = 0;
          more = (cf + 1 < nchan);
        }
        if (more) {
          const CohSegment sn = a.segs[nxt];
          const unsigned bytes = (unsigned)sn.count * (unsigned)sizeof(DevSource);
          mbar_expect_tx(&bar[b ^ 1], bytes);
          if (bytes) bulk_g2s(&sbuf[b ^ 1][0], a.src + sn.first, bytes, &bar[b ^ 1]);
        }
      }
      const CohSegment sg = a.segs[sgi];
      mbar_wait(&bar[b], (unsigned)((it >> 1) & 1));
      if (active) {
        for (int s = 0; s < sg.count; s++) {
          const DevSource &S = sbuf[b][s];
          double2 ph = source_phase(S, a.modes, u, v, w, freq, a.fdelta2);
          // station beams towards this source at this timeslot and channel
          const size_t bt = ((size_t)tslot * nchan + cf) * a.beam_S + (size_t)(sg.first + s);
          if (a.beam_af) {  // array factor of both stations (predict_withbeam.c:336-343)
            const double af = a.beam_af[bt * a.N + s1] * a.beam_af[bt * a.N + s2];
            ph.x *= af;
            ph.y *= af;
          }
          double I = S.sI, Q = S.sQ, U = S.sU, V = S.sV;
          if (MODE >= 1 && S.spec_idx != 0.0) {
            const double fr = log(freq / S.f0);
            const double fr1 = fr * fr, fr2 = fr1 * fr;
            const double tf = S.spec_idx * fr + S.spec_idx1 * fr1 + S.spec_idx2 * fr2;
            I = spec_flux(S.sI0, tf);
            Q = spec_flux(S.sQ0, tf);
            U = spec_flux(S.sU0, tf);
            V = spec_flux(S.sV0, tf);
          }
          if (a.beam_E) {  // E_p (Stokes coherency) E_q^H  (predict_withbeam.c:381-404)
            double2 C0[4], T1[4], E1[4], E2[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
              C0[c] = make_double2(0.0, 0.0);
              E1[c] = a.beam_E[(bt * a.N + s1) * 4 + c];
              E2[c] = a.beam_E[(bt * a.N + s2) * 4 + c];
            }
            add_stokes(C0, ph, I, Q, U, V);
            mat_ab(E1, C0, T1);
            mat_abh(T1, E2, C0);
#pragma unroll
            for (int c = 0; c < 4; c++) C[c] = cadd(C[c], C0[c]);
          } else {
            add_stokes(C, ph, I, Q, U, V);
          }
        }
        if (sg.last) {
          if (MODE == 0) {
            double2 *ck = a.coh + (long long)sg.cluster * 4 * a.R;
#pragma unroll
            for (int c = 0; c < 4; c++) st_stream(ck + (long long)c * a.R + r, C[c]);
          } else if (MODE == 1) {
#pragma unroll
            for (int c = 0; c < 4; c++) X[c] = cadd(X[c], C[c]);
          } else if (a.clus_sub[sg.cluster]) {
            // Jones of this row's hybrid chunk: px = row / ceil(R / nchunk)  (residual.c:717)
            const int nch = a.clus_nchunk[sg.cluster];
            const int px = row_chunk(r, a.R, nch);
            const double *pm = a.p + a.chunk_poff[a.clus_chunk0[sg.cluster] + px];
            double2 G1[4], G2[4], T1[4], T2[4];
            load_jones(pm, s1, G1);
            load_jones(pm, s2, G2);
            mat_ab(G1, C, T1);
            mat_abh(T1, G2, T2);
#pragma unroll
            for (int c = 0; c < 4; c++) X[c] = csub(X[c], T2[c]);
          }
#pragma unroll
          for (int c = 0; c < 4; c++) C[c] = make_double2(0.0, 0.0);
        }
      }
      __syncthreads();  // everyone is done with sbuf[b] before it is refilled two iterations on
    }
    if (MODE >= 1 && active) {
      double2 *xo = a.xout + ((long long)cf * a.R + r) * 4;
      double2 V[4];
#pragma unroll
      for (int c = 0; c < 4; c++) V[c] = cadd(xo[c], X[c]);
      if (MODE == 2 && a.pinv) {
        const int px = row_chunk(r, a.R, a.pinv_nchunk);
        const double *pm = a.pinv + (size_t)8 * a.N * px;
        double2 G1[4], G2[4], T1[4];
        load_jones(pm, s1, G1);
        load_jones(pm, s2, G2);
        mat_ab(G1, V, T1);
        mat_abh(T1, G2, V);
      }
#pragma unroll
      for (int c = 0; c < 4; c++) xo[c] = V[c];
    }
  }
  if (MODE == 0 && active && a.flag) {
    // uv cut: unflagged rows outside [uvmin, uvmax] get flag 2 (predict.c:488-493)
    if (a.flag[r] == 0) {
      const double uvdist = sqrt(u * u + v * v) * a.freqs[0];
      if (uvdist < a.uvmin || uvdist > a.uvmax) a.flag[r] = 2;
    }
  }
}

// ---- station beam tables ----------------------------------------------------------------------------
// JD -> Greenwich mean sidereal angle in degrees (jd2gmst, transforms.c:139-146)
__device__ __forceinline__ double jd2gmst_deg(double time_jd) {
  const double t = (time_jd - 2451545.0) / 36525.0;
  const double theta =
      67310.54841 + t * ((876600.0 * 3600.0 + 8640184.812866) + t * (0.093104 - (6.2 * 10e-6) * t));
  return fmod(fmod(theta, 86400.0 * (theta / fabs(theta))) / 240.0, 360.0);
}
// (ra, dec) -> (az, el) at a station (radec2azel_gmst, transforms.c:157-180)
__device__ __forceinline__ void radec2azel(double ra, double dec, double lon, double lat,
                                           double gmst, double *az, double *el) {
  const double lst = gmst + lon * 180.0 * M_1_PI;
  const double LHA = fmod(lst - ra * 180.0 * M_1_PI, 360.0);
  double sinlat, coslat, sindec, cosdec, sinL, cosL;
  sincos(lat, &sinlat, &coslat);
  sincos(dec, &sindec, &cosdec);
  sincos(LHA * M_PI / 180.0, &sinL, &cosL);
  const double tmp = sinlat * sindec + coslat * cosdec * cosL;
  *el = asin(tmp);
  double sinel, cosel;
  sincos(*el, &sinel, &cosel);
  double a = fmod(atan2(-sinL * cosdec / cosel, (sindec - sinel * sinlat) / (cosel * coslat)),
                  2.0 * M_PI);
  if (a < 0) a += 2.0 * M_PI;
  *az = a;
}
// generalised Laguerre polynomial L_p^q(x) (L_g1, elementbeam.c:341-356)
__device__ __forceinline__ double laguerre(int p, int q, double x) {
  if (p == 0) return 1.0;
  if (p == 1) return 1.0 - x + (double)q;
  double Lp = 0.0, Lp1 = 1.0 - x + (double)q, Lp2 = 1.0;
  for (int i = 2; i <= p; i++) {
    const double p1 = 1.0 / (double)i;
    Lp = (2.0 + p1 * ((double)q - 1.0 - x)) * Lp1 - (1.0 + p1 * (q - 1)) * Lp2;
    Lp2 = Lp1;
    Lp1 = Lp;
  }
  return Lp;
}
// element pattern (theta, phi components) at zenith angle r and azimuth th (eval_elementcoeffs[_wb],
// elementbeam.c:384-460); coefficient set `fi` of the wide-band tables
__device__ __forceinline__ void element_eval(const BeamArgs &a, double r, double th, int fi,
                                             double2 *e_theta, double2 *e_phi) {
  const double rb = pow(r / a.ecbeta, 2);
  const double ex = exp(-0.5 * rb);
  double2 ph = make_double2(0.0, 0.0), tt = make_double2(0.0, 0.0);
  int idx = 0;
  for (int n = 0; n < a.ecM; n++)
    for (int m = -n; m <= n; m += 2) {
      const int absm = m >= 0 ? m : -m;
      const double Lg = laguerre((n - absm) / 2, absm, rb);
      const double rm = pow(M_PI_4 + r, (double)absm);
      double s, c;
      sincos(-(double)m * th, &s, &c);
      const double pr = rm * Lg * ex * a.preamble[idx];
      const double2 basis = make_double2(pr * c, pr * s);
      cfma(ph, a.pat_phi[(size_t)fi * a.ecNmodes + idx], basis);
      cfma(tt, a.pat_theta[(size_t)fi * a.ecNmodes + idx], basis);
      idx++;
    }
  *e_theta = tt;
  *e_phi = ph;
}

// one thread per (timeslot, channel, source, station): array factor (arraybeam / array_element_beam,
// stationbeam.c:49-330) and element E-Jones (element_beam, :372-430)
__global__ void __launch_bounds__(128) k_beam_tables(BeamArgs a) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)a.T * a.Nf * a.S * a.N;
  if (gid >= total) return;
  const int sta = (int)(gid % a.N);
  size_t q = gid / a.N;
  const int s = (int)(q % a.S);
  q /= a.S;
  const int cf = (int)(q % a.Nf);
  const int t = (int)(q / a.Nf);
  const double gmst = jd2gmst_deg(a.time_jd[t]);
  const double ra = a.src[s].ra, dec = a.src[s].dec;
  const double f = a.freqs[cf];
  double az, el;
  radec2azel(ra, dec, a.lon[sta], a.lat[sta], gmst, &az, &el);
  const double theta = M_PI_2 - el;
  if (a.af) {
    double gain = 0.0;
    if (el >= 0.0) {
      const double tpc = 2.0 * M_PI / 299792458.0;
      const double beam_f = a.wideband ? f : a.f0;
      double az0, el0;
      radec2azel(a.ra0, a.dec0, a.lon[sta], a.lat[sta], gmst, &az0, &el0);
      double sint, cost, sinph, cosph, sint0, cost0, sinph0, cosph0;
      sincos(theta, &sint, &cost);
      sincos(-az, &sinph, &cosph);
      sincos(M_PI_2 - el0, &sint0, &cost0);
      sincos(-az0, &sinph0, &cosph0);
      double rat1 = beam_f * sint0;
      const double rat2 = f * sint;
      double r1 = rat1 * cosph0 - rat2 * cosph, r2 = rat1 * sinph0 - rat2 * sinph;
      double r3 = beam_f * cost0 - f * cost;
      const int K = a.Nelem[sta];
      const double *px = a.ex + a.elem_off[sta], *py = a.ey + a.elem_off[sta];
      const double *pz = a.ez + a.elem_off[sta];
      const int skip = a.bf_type == 2 ? 16 : 0;  // STAT_TILE: tile centroids follow the 16 dipoles
      double csum = 0.0, ssum = 0.0;
      for (int j = 0; j < K; j++) {
        double sn, cs;
        sincos(-tpc * (r1 * px[j + skip] + r2 * py[j + skip] + r3 * pz[j + skip]), &sn, &cs);
        ssum += sn;
        csum += cs;
      }
      if (a.bf_type == 2) {
        double azb, elb;
        radec2azel(a.b_ra0, a.b_dec0, a.lon[sta], a.lat[sta], gmst, &azb, &elb);
        sincos(M_PI_2 - elb, &sint0, &cost0);
        sincos(-azb, &sinph0, &cosph0);
        rat1 = beam_f * sint0;
        r1 = rat1 * cosph0 - rat2 * cosph;
        r2 = rat1 * sinph0 - rat2 * sinph;
        r3 = beam_f * cost0 - f * cost;
        double cb = 0.0, sb = 0.0;
        for (int j = 0; j < 16; j++) {
          double sn, cs;
          sincos(-tpc * (r1 * px[j] + r2 * py[j] + r3 * pz[j]), &sn, &cs);
          sb += sn;
          cb += cs;
        }
        gain = sqrt(csum * csum + ssum * ssum) * sqrt(cb * cb + sb * sb) / (double)(K * 16);
      } else {
        gain = sqrt(csum * csum + ssum * ssum) / (double)K;
      }
    }
    a.af[gid] = gain;
  }
  if (a.E) {
    double2 e[4];
    e[0] = e[1] = e[2] = e[3] = make_double2(0.0, 0.0);
    if (el >= 0.0) {
      const int fi = a.wideband ? cf : 0;
      // E = [E_theta(az - pi/4) E_phi(az - pi/4); E_theta(az + pi/4) E_phi(az + pi/4)]
      element_eval(a, theta, az - M_PI_4, fi, &e[0], &e[1]);
      element_eval(a, theta, az - M_PI_4 + M_PI_2, fi, &e[2], &e[3]);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) a.E[gid * 4 + c] = e[c];
  }
}

extern "C" {
void db_launch_beam_tables(const BeamArgs *a, cudaStream_t st) {
  const size_t total = (size_t)a->T * a->Nf * a->S * a->N;
  k_beam_tables<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(*a);
}
void db_launch_coherencies(const CohArgs *a, cudaStream_t st) {
  unsigned grid = (unsigned)((a->R + COH_THREADS - 1) / COH_THREADS);
  k_sky_predict<0><<<grid, COH_THREADS, 0, st>>>(*a);
}
void db_launch_residual_multifreq(const CohArgs *a, cudaStream_t st) {
  unsigned grid = (unsigned)((a->R + COH_THREADS - 1) / COH_THREADS);
  k_sky_predict<2><<<grid, COH_THREADS, 0, st>>>(*a);
}
void db_launch_predict_multifreq(const CohArgs *a, cudaStream_t st) {
  unsigned grid = (unsigned)((a->R + COH_THREADS - 1) / COH_THREADS);
  k_sky_predict<1><<<grid, COH_THREADS, 0, st>>>(*a);
}
}
