"""Synthetic measurement-set generator for the BASELINE.json configs.

Everything is produced in memory with numpy in the layouts the Dirac C API takes:
  u,v,w   [Nbase*tilesz]            seconds (= metres / c, fullbatch_mode.cpp:320-322)
  x       [8*Nbase*tilesz]          XX(re,im) XY YX YY per row, row = t*Nbase + b
  coh     [4*M*Nbase*tilesz] c128   coh[4*M*row + 4*k + c]  (predict.c:484-487)
  pp      [8*N*Mt]                  Jones, row-major 2x2, (re,im) pairs (lmfit.c:90-97)
Row order is the one `generate_baselines` produces (baseline_utils.c:438-466).

No reference code is used here; coherencies of point (and Gaussian) sources are computed with
the published formula (phase * |sinc| frequency smearing * Stokes, predict.c:345-497) so that
the generator also works on the GPU box where /root/reference does not exist.
"""
from __future__ import annotations

import dataclasses
import numpy as np

C_LIGHT = 299792458.0

CONFIGS = {
    # name: N, M, tilesz, disc radius (m), solver settings
    "C1": dict(N=8, M=2, tilesz=10, radius=1.5e3, seed=20260921 + 1, max_iter=5, kmean=0.0),
    "C2": dict(N=62, M=64, tilesz=120, radius=40e3, seed=20260921 + 2, max_iter=2, kmean=2.0),
    "C3": dict(N=62, M=64, tilesz=120, radius=40e3, seed=20260921 + 3, max_iter=2, kmean=2.0,
               outliers=0.02),
    "C4": dict(N=512, M=256, tilesz=120, radius=75e3, seed=20260921 + 4, max_iter=2, kmean=2.0),
    "C5": dict(N=62, M=128, tilesz=120, radius=40e3, seed=20260921 + 5, max_iter=2, kmean=2.0),
}


@dataclasses.dataclass
class Problem:
    N: int
    Nbase: int
    tilesz: int
    M: int
    Mt: int
    freq0: float
    fdelta: float
    u: np.ndarray
    v: np.ndarray
    w: np.ndarray
    sta1: np.ndarray
    sta2: np.ndarray
    flag: np.ndarray          # uint8 per row: 0 ok, 1 flagged, 2 uv-cut
    clusters: list            # list of dicts (see dirac_api.SkyModel)
    coh: np.ndarray | None    # complex128 [Nbase1*M*4]
    x: np.ndarray | None      # float64 [8*Nbase1]
    jones_true: np.ndarray    # float64 [8*N*Mt]
    pp0: np.ndarray           # identity initial Jones [8*N*Mt]
    nchunk: list

    @property
    def Nbase1(self):
        return self.Nbase * self.tilesz


def baseline_pairs(N: int):
    """(sta1, sta2) in the canonical order (0,1),(0,2)...(N-2,N-1), baseline_utils.c:445-461."""
    p, q = np.triu_indices(N, k=1)
    return p.astype(np.int32), q.astype(np.int32)


def station_uvw(N, radius, tilesz, rng, dec0=np.deg2rad(60.0), ha0=-np.pi / 12.0, dt=10.0):
    """Per-station (u,v,w) tracks by Earth-rotation synthesis, metres."""
    r = radius * np.sqrt(rng.uniform(0, 1, N))
    th = rng.uniform(0, 2 * np.pi, N)
    X = r * np.cos(th)
    Y = r * np.sin(th)
    Z = rng.normal(0, 10.0, N)
    ha = ha0 + (2 * np.pi / 86164.0905) * dt * np.arange(tilesz)
    sh, ch = np.sin(ha)[:, None], np.cos(ha)[:, None]
    sd, cd = np.sin(dec0), np.cos(dec0)
    us = sh * X + ch * Y
    vs = -sd * ch * X + sd * sh * Y + cd * Z
    ws = cd * ch * X - cd * sh * Y + sd * Z
    return us, vs, ws  # [tilesz, N]


def make_clusters(M, N, rng, kmean=2.0, nchunk=None, fov_deg=5.0, gaussian_frac=0.0):
    clusters = []
    for k in range(M):
        K = 1 + (rng.poisson(kmean) if kmean > 0 else 0)
        rr = np.deg2rad(fov_deg) * np.sqrt(rng.uniform(0, 1))
        ang = rng.uniform(0, 2 * np.pi)
        l0, m0 = rr * np.cos(ang), rr * np.sin(ang)
        dl = np.deg2rad(0.2) * rng.uniform(-1, 1, K)
        dm = np.deg2rad(0.2) * rng.uniform(-1, 1, K)
        dl[0] = 0.0
        dm[0] = 0.0
        ll = l0 + dl
        mm = m0 + dm
        nn = np.sqrt(1.0 - ll * ll - mm * mm) - 1.0
        sI = rng.lognormal(0.0, 1.0, K)
        cl = dict(ll=ll, mm=mm, nn=nn, sI=sI, sQ=np.zeros(K), sU=np.zeros(K), sV=np.zeros(K),
                  stype=np.zeros(K, dtype=np.uint8), nchunk=1 if nchunk is None else int(nchunk[k]),
                  id=k)
        if gaussian_frac > 0:
            g = rng.uniform(0, 1, K) < gaussian_frac
            cl["stype"] = g.astype(np.uint8)  # STYPE_GAUSSIAN == 1
            gauss = np.zeros((K, 8))
            gauss[:, 0] = np.deg2rad(rng.uniform(0.5, 3.0, K) / 60.0)   # eX (rad)
            gauss[:, 1] = np.deg2rad(rng.uniform(0.5, 3.0, K) / 60.0)   # eY
            gauss[:, 2] = rng.uniform(0, np.pi, K)                      # eP
            gauss[:, 3] = 1.0  # cxi
            gauss[:, 5] = 1.0  # cphi
            gauss[:, 7] = 0    # use_projection
            cl["gauss"] = gauss
            cl["sQ"] = 0.1 * sI * rng.uniform(-1, 1, K)
            cl["sU"] = 0.1 * sI * rng.uniform(-1, 1, K)
            cl["sV"] = 0.02 * sI * rng.uniform(-1, 1, K)
        clusters.append(cl)
    return clusters


def coherencies(u, v, w, clusters, freq0, fdelta, rows=None):
    """coh[row][k][4] for point/Gaussian sources; restates predict.c:411-487 in numpy."""
    if rows is not None:
        u, v, w = u[rows], v[rows], w[rows]
    nrow = len(u)
    M = len(clusters)
    coh = np.zeros((nrow, M, 4), dtype=np.complex128)
    for k, cl in enumerate(clusters):
        G = 2.0 * np.pi * (u[:, None] * cl["ll"][None, :] + v[:, None] * cl["mm"][None, :]
                           + w[:, None] * cl["nn"][None, :])
        ph = np.exp(1j * G * freq0)
        sm = G * (0.5 * fdelta)
        with np.errstate(invalid="ignore", divide="ignore"):
            fac = np.where(G != 0.0, np.abs(np.sin(sm) / sm), 1.0)
        ph = ph * fac
        if "gauss" in cl:
            g = cl["gauss"]
            uu = u[:, None] * freq0
            vv = v[:, None] * freq0
            sp, cp = np.sin(g[:, 2])[None, :], np.cos(g[:, 2])[None, :]
            ut = g[:, 0][None, :] * (cp * uu - sp * vv)
            vt = g[:, 1][None, :] * (sp * uu + cp * vv)
            shape = np.exp(-2.0 * np.pi ** 2 * (ut * ut + vt * vt))
            ph = np.where((cl["stype"] == 1)[None, :], ph * shape, ph)
        I, Q, U, V = cl["sI"], cl["sQ"], cl["sU"], cl["sV"]
        coh[:, k, 0] = ph @ (I + Q)
        coh[:, k, 1] = ph @ (U + 1j * V)
        coh[:, k, 2] = ph @ (U - 1j * V)
        coh[:, k, 3] = ph @ (I - Q)
    return coh.reshape(-1)


def chunk_index(rows, Nbase1, nchunk):
    """px = row / ceil(Nbase1/nchunk)  (lmfit.c:86,655)."""
    return rows // ((Nbase1 + nchunk - 1) // nchunk)


def apply_jones(coh, pp, sta1, sta2, N, clusters_nchunk, flag=None):
    """Model visibilities sum_k J_p C J_q^H as float64 [8*Nbase1]; restates lmfit.c:611-688."""
    M = len(clusters_nchunk)
    nrow = len(sta1)
    c = coh.reshape(nrow, M, 2, 2)
    out = np.zeros((nrow, 2, 2), dtype=np.complex128)
    rows = np.arange(nrow)
    off = 0
    for k in range(M):
        nch = clusters_nchunk[k]
        px = chunk_index(rows, nrow, nch)
        J = pp[off:off + nch * 8 * N].reshape(nch, N, 4, 2)
        J = (J[..., 0] + 1j * J[..., 1]).reshape(nch, N, 2, 2)
        Jp = J[px, sta1]
        Jq = J[px, sta2]
        out += Jp @ c[:, k] @ np.conj(np.swapaxes(Jq, 1, 2))
        off += nch * 8 * N
    if flag is not None:
        out[flag != 0] = 0.0
    o = np.empty((nrow, 4, 2))
    o[:, :, 0] = out.reshape(nrow, 4).real
    o[:, :, 1] = out.reshape(nrow, 4).imag
    return o.reshape(-1)


def make_problem(N, M, tilesz, radius=1.5e3, seed=1, kmean=0.0, freq0=150e6, fdelta=195.3e3,
                 nchunk=None, flag_frac=0.01, uvcut_frac=0.005, noise_rel=1e-2, jones_amp=0.2,
                 outliers=0.0, gaussian_frac=0.0, with_data=True, **_unused) -> Problem:
    rng = np.random.default_rng(seed)
    us, vs, ws = station_uvw(N, radius, tilesz, rng)
    p, q = baseline_pairs(N)
    Nbase = len(p)
    u = ((us[:, p] - us[:, q]) / C_LIGHT).reshape(-1)
    v = ((vs[:, p] - vs[:, q]) / C_LIGHT).reshape(-1)
    w = ((ws[:, p] - ws[:, q]) / C_LIGHT).reshape(-1)
    sta1 = np.tile(p, tilesz)
    sta2 = np.tile(q, tilesz)
    Nbase1 = Nbase * tilesz
    clusters = make_clusters(M, N, rng, kmean=kmean, nchunk=nchunk, gaussian_frac=gaussian_frac)
    nch = [cl["nchunk"] for cl in clusters]
    Mt = int(sum(nch))
    # flags: 1 % flagged rows, uv-cut on the shortest 0.5 %
    flag = (rng.uniform(0, 1, Nbase1) < flag_frac).astype(np.uint8)
    if uvcut_frac > 0:
        uvd = np.sqrt(u * u + v * v)
        cut = np.quantile(uvd, uvcut_frac)
        flag[(uvd < cut) & (flag == 0)] = 2
    jt = np.zeros((Mt, N, 4, 2))
    jt[:, :, 0, 0] = 1.0
    jt[:, :, 3, 0] = 1.0
    jt += jones_amp * rng.normal(0, 1, jt.shape) / np.sqrt(2.0)
    jones_true = jt.reshape(-1)
    pp0 = np.zeros((Mt, N, 8))
    pp0[:, :, 0] = 1.0
    pp0[:, :, 6] = 1.0
    pp0 = pp0.reshape(-1)
    coh = x = None
    if with_data:
        coh = coherencies(u, v, w, clusters, freq0, fdelta)
        x = apply_jones(coh, jones_true, sta1, sta2, N, nch)
        sigma = noise_rel * np.median(np.abs(x))
        noise = rng.normal(0, sigma, x.shape)
        if outliers > 0:
            bad = rng.uniform(0, 1, x.shape) < outliers
            noise[bad] += 20.0 * sigma * rng.choice([-1.0, 1.0], size=int(bad.sum()))
        x = x + noise
        x.reshape(Nbase1, 8)[flag == 1] = 0.0   # preset_flags_and_data, baseline_utils.c:206-227
    return Problem(N=N, Nbase=Nbase, tilesz=tilesz, M=M, Mt=Mt, freq0=freq0, fdelta=fdelta,
                   u=u, v=v, w=w, sta1=sta1, sta2=sta2, flag=flag, clusters=clusters, coh=coh,
                   x=x, jones_true=jones_true, pp0=pp0, nchunk=nch)


def make_config(name: str, **over) -> Problem:
    cfg = dict(CONFIGS[name])
    cfg.update(over)
    return make_problem(**cfg)
