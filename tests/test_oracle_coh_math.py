"""CPU: the per-source / per-station arithmetic of the product's coherency kernels
(sagecal_b200/csrc/coh_math.cuh, compiled as host code by oracle/coh_math_check.cu) against the
compiled reference: shapelet_contrib (shapelet.c:141), arraybeam (stationbeam.c:49),
array_element_beam (:191) and element_beam (:372).  Runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from sagecal_b200.dirac_api import dptr, elementcoeff, exinfo_shapelet

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "oracle", "libcoh_math_check.so")
d, i, dp = C.c_double, C.c_int, C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def chk():
    if not os.path.exists(LIB):
        pytest.skip("oracle/libcoh_math_check.so not built (make -C oracle)")
    L = C.CDLL(LIB)
    L.check_shapelet.argtypes = [i, d, dp, d, d, d, d, d, d, d, i, d, d, d, dp]
    L.check_beam.argtypes = [d, d, i, d, d, d, d, d, d, i, dp, dp, d, C.POINTER(i), C.POINTER(i), dp,
                             dp, dp, i, i, i, d, dp, dp, dp, i, dp, dp]
    return L


def test_shapelet_factor(ref, chk):
    rng = np.random.default_rng(3)
    for n0 in (1, 2, 5, 9, 16):
        modes = np.ascontiguousarray(rng.normal(0, 1, n0 * n0) / n0)
        for proj in (0, 1):
            xi, phi = rng.uniform(0, 2 * np.pi), rng.uniform(0, 0.3)
            g = exinfo_shapelet(n0, np.deg2rad(1.5 / 60.0), dptr(modes), rng.uniform(0.7, 1.4),
                                rng.uniform(0.7, 1.4), rng.uniform(0, np.pi), np.cos(xi), np.sin(xi),
                                np.cos(phi), np.sin(phi), proj)
            for _ in range(5):
                u, v, w = rng.normal(0, 800.0, 3)
                out = np.zeros(2)
                chk.check_shapelet(n0, g.beta, dptr(modes), g.eX, g.eY, g.eP, g.cxi, g.sxi, g.cphi,
                                   g.sphi, proj, u, v, w, dptr(out))
                fn = ref.lib.shapelet_contrib
                fn.argtypes = [C.c_void_p, d, d, d]
                fn.restype = _cplx   # complex double by value: two doubles (x86-64 SysV: xmm0, xmm1)
                want = fn(C.byref(g), u, v, w)
                want = complex(want.re, want.im)
                got = complex(out[0], out[1])
                assert abs(got - want) <= 1e-12 * max(1e-300, abs(want)) + 1e-15, (n0, proj, got, want)


class _cplx(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


@pytest.mark.parametrize("tile", [False, True], ids=["single", "tile"])
@pytest.mark.parametrize("wide", [0, 1], ids=["narrow", "wide"])
def test_station_beams(ref, chk, tile, wide):
    rng = np.random.default_rng(11 + tile)
    N = 7
    lon = np.deg2rad(6.87 + rng.uniform(-0.5, 0.5, N))
    lat = np.deg2rad(52.9 + rng.uniform(-0.3, 0.3, N))
    ra0, dec0 = 1.2, np.deg2rad(58.0)
    elems = []
    for n in range(N):
        if tile:
            g = (np.arange(4) - 1.5) * 1.25
            dip = np.array([[x, y, 0.0] for x in g for y in g])
            cen = np.c_[rng.uniform(-15, 15, (18 + n, 2)), rng.normal(0, 0.05, 18 + n)]
            elems.append(np.vstack([dip, cen]))
        else:
            K = 30 + 5 * n
            elems.append(np.c_[rng.uniform(-40, 40, (K, 2)), rng.normal(0, 0.1, K)])
    extra = 16 if tile else 0
    Nelem = np.array([len(e) - extra for e in elems], dtype=np.int32)
    off = np.concatenate([[0], np.cumsum([len(e) for e in elems])[:-1]]).astype(np.int32)
    ex = np.ascontiguousarray(np.concatenate([e[:, 0] for e in elems]))
    ey = np.ascontiguousarray(np.concatenate([e[:, 1] for e in elems]))
    ez = np.ascontiguousarray(np.concatenate([e[:, 2] for e in elems]))
    xs = [np.ascontiguousarray(e[:, 0]) for e in elems]
    ys = [np.ascontiguousarray(e[:, 1]) for e in elems]
    zs = [np.ascontiguousarray(e[:, 2]) for e in elems]
    mk = lambda arrs: (dp * N)(*[dptr(a) for a in arrs])
    xx, yy, zz = mk(xs), mk(ys), mk(zs)
    freqs = np.array([144e6, 151e6])
    ec = elementcoeff()
    if wide:
        ref.lib.set_elementcoeffs_wb(1 if tile else 0, dptr(freqs), len(freqs), C.byref(ec))
    else:
        ref.lib.set_elementcoeffs(1 if tile else 0, d(150e6), C.byref(ec))
    nfc = len(freqs) if wide else 1
    pphi = np.ctypeslib.as_array(C.cast(ec.pattern_phi, dp), shape=(2 * ec.Nmodes * nfc,)).copy()
    pth = np.ctypeslib.as_array(C.cast(ec.pattern_theta, dp), shape=(2 * ec.Nmodes * nfc,)).copy()
    pre = np.ctypeslib.as_array(C.cast(ec.preamble, dp), shape=(ec.Nmodes,)).copy()
    bf = 2 if tile else 1
    f0 = 148e6
    for trial in range(6):
        ra = ra0 + np.deg2rad(rng.uniform(-5, 5))
        dec = dec0 + np.deg2rad(rng.uniform(-5, 5)) if trial < 5 else np.deg2rad(-65.0)  # below horizon
        jd = 2456789.3 + rng.uniform(0, 1)
        findex = int(trial % nfc)
        f = float(freqs[findex])
        want_af, want_E = np.zeros(N), np.zeros(8 * N)
        ref.lib.array_element_beam(d(ra), d(dec), bf, d(ra0 + 0.01), d(dec0 - 0.01), d(ra0), d(dec0),
                                   d(f), d(f0), N, dptr(lon), dptr(lat), d(jd),
                                   Nelem.ctypes.data_as(C.POINTER(i)), xx, yy, zz, C.byref(ec),
                                   dptr(want_af), dptr(want_E), wide, findex)
        only_af, only_E = np.zeros(N), np.zeros(8 * N)
        ref.lib.arraybeam(d(ra), d(dec), bf, d(ra0 + 0.01), d(dec0 - 0.01), d(ra0), d(dec0), d(f),
                          d(f0), N, dptr(lon), dptr(lat), d(jd), Nelem.ctypes.data_as(C.POINTER(i)),
                          xx, yy, zz, dptr(only_af), wide)
        ref.lib.element_beam(d(ra), d(dec), d(f), d(f0), N, dptr(lon), dptr(lat), d(jd), C.byref(ec),
                             dptr(only_E), wide, findex)
        assert np.allclose(only_af, want_af, rtol=1e-13, atol=0) and np.allclose(only_E, want_E, rtol=1e-13)
        got_af, got_E = np.zeros(N), np.zeros(8 * N)
        chk.check_beam(ra, dec, bf, ra0 + 0.01, dec0 - 0.01, ra0, dec0, f, f0, N, dptr(lon), dptr(lat),
                       jd, Nelem.ctypes.data_as(C.POINTER(i)), off.ctypes.data_as(C.POINTER(i)),
                       dptr(ex), dptr(ey), dptr(ez), wide, ec.M, ec.Nmodes, ec.beta, dptr(pphi),
                       dptr(pth), dptr(pre), findex, dptr(got_af), dptr(got_E))
        if trial == 5:
            assert np.all(want_af == 0) and np.all(got_af == 0) and np.all(got_E == 0)
            continue
        assert np.max(np.abs(got_af - want_af)) <= 1e-11 * np.max(np.abs(want_af)), trial
        assert np.max(np.abs(got_E - want_E)) <= 1e-11 * np.max(np.abs(want_E)), trial
