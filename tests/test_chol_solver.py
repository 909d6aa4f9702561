"""Cluster Cholesky solver (kernels_chol.cu) against numpy on SPD systems of the sizes the LM uses."""
import ctypes as C

import numpy as np
import pytest

from sagecal_b200 import lib as blib

pytestmark = pytest.mark.gpu


def _solve(n, A, b, mu):
    api = blib.load()
    L = api.lib
    L.dirac_b200_spd_solve.restype = C.c_int
    L.dirac_b200_spd_solve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    x = np.zeros(n)
    info = np.zeros(2, dtype=np.int32)
    A = np.asfortranarray(A)
    rc = L.dirac_b200_spd_solve(n, A.ctypes.data, b.ctypes.data, mu, x.ctypes.data, info.ctypes.data)
    return rc, x, int(info[0])


@pytest.mark.parametrize("n", [8, 31, 32, 33, 64, 200, 496, 512])
def test_spd_solve_matches_numpy(n):
    rng = np.random.default_rng(n)
    J = rng.standard_normal((2 * n, n))
    A = J.T @ J
    b = rng.standard_normal(n)
    mu = 1e-3 * np.max(np.diag(A))
    rc, x, info = _solve(n, A, b, mu)
    assert rc == 0 and info == 0
    ref = np.linalg.solve(A + mu * np.eye(n), b)
    # tolerance: backward-stable factorisation, cond ~1e3..1e4
    assert np.max(np.abs(x - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_spd_solve_only_reads_lower_triangle():
    n = 100
    rng = np.random.default_rng(1)
    J = rng.standard_normal((3 * n, n))
    A = J.T @ J
    b = rng.standard_normal(n)
    Al = np.tril(A) + np.triu(np.full((n, n), np.nan), 1)
    rc, x, info = _solve(n, Al, b, 0.5)
    assert rc == 0 and info == 0
    ref = np.linalg.solve(A + 0.5 * np.eye(n), b)
    assert np.allclose(x, ref, rtol=1e-10, atol=1e-12)


def test_spd_solve_reports_failing_pivot():
    n = 96
    A = np.eye(n)
    A[40, 40] = -1.0
    b = np.ones(n)
    rc, x, info = _solve(n, A, b, 0.0)
    assert rc == 0 and info == 41


@pytest.mark.parametrize("n", [16, 33, 250, 496, 512])
def test_tri_solve_matches_numpy(n):
    """Solve-only cluster kernel on a factor laid out as LAPACK/cuSOLVER leave it (lower, ld = n)."""
    api = blib.load()
    L = api.lib
    L.dirac_b200_tri_solve.restype = C.c_int
    L.dirac_b200_tri_solve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    rng = np.random.default_rng(100 + n)
    J = rng.standard_normal((2 * n, n))
    A = J.T @ J + 0.1 * np.eye(n)
    fac = np.linalg.cholesky(A)
    # upper triangle poisoned: it must never be read
    Lf = np.asfortranarray(np.tril(fac) + np.triu(np.full((n, n), np.nan), 1))
    b = rng.standard_normal(n)
    x = np.zeros(n)
    rc = L.dirac_b200_tri_solve(n, Lf.ctypes.data, b.ctypes.data, x.ctypes.data, 0, None)
    if rc == -1:
        pytest.skip("cluster size on this device too small for the solve-only kernel")
    ref = np.linalg.solve(A, b)
    assert np.max(np.abs(x - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_pivot_rsqrt_accuracy():
    """The branch-free 1/sqrt of the pivot chain (hardware seed + one cubic correction)."""
    api = blib.load()
    L = api.lib
    L.dirac_b200_test_rsqrt.restype = C.c_int
    L.dirac_b200_test_rsqrt.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(7)
    x = np.concatenate([10.0 ** rng.uniform(-30, 30, 200000), rng.uniform(0.5, 4.0, 200000)])
    y = np.zeros_like(x)
    L.dirac_b200_test_rsqrt(x.size, x.ctypes.data, y.ctypes.data)
    ref = 1.0 / np.sqrt(x.astype(np.longdouble))
    rel = np.max(np.abs((y - ref) / ref).astype(np.float64))
    assert rel <= 4 * np.finfo(np.float64).eps, rel


@pytest.mark.parametrize("n", [576, 1024, 4096])
def test_bigtri_solve_matches_numpy(api, n):
    """blocked dataflow substitutions for systems beyond the cluster kernels (8N > 512) against
    scipy on the same factor"""
    import ctypes as C
    import scipy.linalg as sla
    from sagecal_b200.dirac_api import dptr
    rng = np.random.default_rng(n)
    A = rng.normal(0, 1, (n, n))
    A = A @ A.T / n + np.eye(n) * 0.5
    Lf = np.linalg.cholesky(A)
    b = rng.normal(0, 1, n)
    want = sla.cho_solve((Lf, True), b)
    Lcol = np.asfortranarray(Lf)          # column-major lower, ld = n
    x = np.zeros(n)
    us = C.c_double(0.0)
    api.lib.dirac_b200_bigtri_solve.restype = C.c_int
    rc = api.lib.dirac_b200_bigtri_solve(n, Lcol.ctypes.data_as(C.POINTER(C.c_double)), dptr(b), dptr(x), 20,
                                         C.byref(us))
    assert rc == 0
    assert np.max(np.abs(x - want)) <= 1e-10 * np.max(np.abs(want))
    print("bigtri n=%d: %.1f us per solve" % (n, us.value))
