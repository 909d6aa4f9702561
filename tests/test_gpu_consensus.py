"""GPU: consensus (ADMM) entry points.
 - dirac_b200_consensus_step against its numpy restatement (single subband: the all-reduce is the
   identity);
 - the consensus-augmented LM (dirac_b200_sagefit_admm / sagefit_visibilities_admm): with rho = 0, Y = 0
   it is the plain SAGE + LM solve; with rho > 0 every sweep is a coordinate descent on
   ||x - f(J)||^2 + Y^T (J - BZ) + rho/2 |J - BZ|^2, so that cost decreases, and a very large rho pins
   J to BZ."""
import ctypes as C

import numpy as np
import pytest

from util import small_problem, relerr
from sagecal_b200 import consensus as cons
from sagecal_b200 import lib as blib
from sagecal_b200.dirac_api import dptr

pytestmark = pytest.mark.gpu


def aug_cost(dp, pp, Y, BZ, rho_i):
    return dp.cost(pp) + float(np.sum(Y * (pp - BZ)) + 0.5 * np.sum(rho_i * (pp - BZ) ** 2))


def test_consensus_step_matches_numpy(api):
    b = small_problem(N=9, M=4, tilesz=4, seed=3, nchunk=[1, 2, 1, 1])
    pr = b.pr
    rng = np.random.default_rng(1)
    m = b.m
    J, Y0, BZ0 = rng.normal(0, 1, m), rng.normal(0, 0.1, m), rng.normal(0, 1, m)
    rho = np.array([3.0, 7.0, 0.5, 12.0])
    freqs = np.array([150e6])
    with blib.DeviceProblem(api, pr.N, pr.Nbase, pr.tilesz, b.barr, b.sky, pr.coh, pr.x) as dp:
        sb = cons.ConsensusSubband(api, dp, 0, freqs, 150e6, 1, rho, ptype=0)
        sb.Y, sb.BZ = Y0.copy(), BZ0.copy()
        pri, dua = sb.exchange(J)
        clus_of = np.concatenate([np.full(8 * pr.N * n, k) for k, n in enumerate(pr.nchunk)])
        Yn, bz, p2, d2 = cons.step_numpy(J, Y0, BZ0, rho[clus_of], sb.Bf, sb.Bi, clus_of, lambda z: None)
        # (one subband, one basis function: Y + rho (J - BZ) cancels to rounding level)
        assert relerr(sb.BZ, bz) < 1e-14
        assert np.max(np.abs(sb.Y - Yn)) < 1e-12 * np.max(np.abs(Y0 + rho[clus_of] * J))
        assert abs(pri - p2) <= 1e-12 * p2 and abs(dua - d2) <= 1e-12 * d2


@pytest.fixture
def admm_lm(api):
    """the drop-in entry point with this library's LM on the augmented cost (default: robust RTR)"""
    api.set_option("admm_lm", 1)
    yield
    api.set_option("admm_lm", 0)


ADMM_CASES = [
    ("admm", dict(N=10, M=3, tilesz=8, seed=6, kmean=1.0, outliers=0.02), dict(max_iter=2)),
    ("admm-hybrid", dict(N=12, M=4, tilesz=10, seed=7, nchunk=[1, 2, 1, 5], outliers=0.02),
     dict(max_iter=3)),
    ("admm-62", dict(N=62, M=2, tilesz=3, seed=8), dict(max_iter=2, max_emiter=2)),
]


@pytest.mark.parametrize("name,prob,args", ADMM_CASES, ids=[c[0] for c in ADMM_CASES])
def test_sagefit_admm_matches_reference(api, refser, name, prob, args):
    """sagefit_visibilities_admm (admm_solve.c:221-420: every visit by rtr_solve_nocuda_robust_admm)
    against the compiled reference (serialised-thread build: the nu update of the threaded one
    races)"""
    b = small_problem(**prob)
    pr = b.pr
    rng = np.random.default_rng(11)
    BZ = pr.jones_true + 0.03 * rng.normal(0, 1, b.m)
    Y = 0.2 * rng.normal(0, 1, b.m)
    rho = rng.uniform(2.0, 30.0, pr.M)
    kw = dict(max_emiter=3, max_iter=2)
    kw.update(args)
    out = []
    for lib in (refser, api):
        x, pp = pr.x.copy(), pr.pp0.copy()
        r = lib.sagefit_visibilities_admm(pr.u, pr.v, pr.w, x, pr.N, pr.Nbase, pr.tilesz,
                                          b.fresh_barr(), b.sky, pr.coh, pp, Y.copy(), BZ.copy(), rho,
                                          **kw)
        out.append((r, x, pp))
    (rr, xr, ppr), (rg, xg, ppg) = out
    assert rr[0] == rg[0]
    assert abs(rr[1] - rg[1]) < 1e-9                    # mean nu
    assert abs(rr[2] - rg[2]) <= 1e-10 * rr[2]
    assert relerr(ppg, ppr) < 1e-5, (name, relerr(ppg, ppr))
    assert relerr(xg, xr) < 1e-5
    assert abs(rr[3] - rg[3]) <= 1e-5 * rr[3]


def test_admm_with_zero_rho_is_the_plain_solve(api, admm_lm):
    b = small_problem(N=10, M=3, tilesz=6, seed=4, kmean=1.0)
    pr = b.pr
    Y = np.zeros(b.m)
    BZ = np.zeros(b.m)
    rho = np.zeros(pr.M)
    api.lib.sagefit_visibilities_admm.restype = C.c_int
    x1, p1 = pr.x.copy(), pr.pp0.copy()
    r1 = api.sagefit_visibilities(pr.u, pr.v, pr.w, x1, pr.N, pr.Nbase, pr.tilesz, b.fresh_barr(), b.sky,
                                  pr.coh, p1, max_emiter=3, max_iter=3, max_lbfgs=0, solver_mode=1)
    x2, p2 = pr.x.copy(), pr.pp0.copy()
    nu, r0, rr = C.c_double(0), C.c_double(0), C.c_double(0)
    from sagecal_b200.dirac_api import cptr
    rv = api.lib.sagefit_visibilities_admm(
        dptr(pr.u), dptr(pr.v), dptr(pr.w), dptr(x2), pr.N, pr.Nbase, pr.tilesz, b.fresh_barr(),
        b.sky.arr, cptr(pr.coh), b.sky.M, b.sky.Mt, C.c_double(150e6), C.c_double(195.3e3), dptr(p2),
        dptr(Y), dptr(BZ), C.c_double(0.0), 4, 3, 3, 0, 7, 128, 0, 5, C.c_double(2.0), C.c_double(30.0), 0,
        dptr(rho), C.byref(nu), C.byref(r0), C.byref(rr))
    assert rv == r1[0]
    assert relerr(p2, p1) < 1e-9 and relerr(x2, x1) < 1e-9
    assert abs(rr.value - r1[3]) <= 1e-9 * r1[3]


def test_admm_sweeps_descend_the_augmented_cost(api):
    b = small_problem(N=12, M=4, tilesz=8, seed=5, kmean=1.0)
    pr = b.pr
    rng = np.random.default_rng(9)
    rho = np.array([4.0, 10.0, 1.0, 6.0]) * 50.0
    clus_of = np.repeat(np.arange(pr.M), 8 * pr.N)
    rho_i = rho[clus_of]
    BZ = pr.jones_true + 0.02 * rng.normal(0, 1, b.m)
    Y = 0.5 * rng.normal(0, 1, b.m)
    with blib.DeviceProblem(api, pr.N, pr.Nbase, pr.tilesz, b.barr, b.sky, pr.coh, pr.x) as dp:
        sb = cons.ConsensusSubband(api, dp, 0, np.array([150e6]), 150e6, 1, rho, ptype=0)
        sb.Y, sb.BZ = Y.copy(), BZ.copy()
        pp = pr.pp0.copy()
        costs = [aug_cost(dp, pp, Y, BZ, rho_i)]
        for _ in range(4):
            rv, r0, r1 = sb.jupdate(pp, max_emiter=1, max_iter=3)
            costs.append(aug_cost(dp, pp, Y, BZ, rho_i))
        assert all(b2 < a2 for a2, b2 in zip(costs, costs[1:])), costs
        assert costs[-1] < 0.2 * costs[0]
        # a very large rho pins J to BZ (minimiser of rho/2 |J - BZ|^2 + y^T (J - BZ): J = BZ - y/rho)
        big = np.full(pr.M, 1e9)
        sb2 = cons.ConsensusSubband(api, dp, 0, np.array([150e6]), 150e6, 1, big, ptype=0)
        sb2.Y, sb2.BZ = Y.copy(), BZ.copy()
        pp2 = pr.pp0.copy()
        for _ in range(3):
            sb2.jupdate(pp2, max_emiter=1, max_iter=4)
        assert np.max(np.abs(pp2 - (BZ - Y / 1e9))) < 1e-5
