"""GPU parity of the multi-channel minibatch (stochastic) robust LBFGS, SURVEY.md 8f-3:
bfgsfit_minibatch_visibilities / bfgsfit_minibatch_consensus with their persistent state, against the
compiled reference (robust_batchmode_lbfgs.c:1446-1577, lbfgs.c:717-930) over several epochs of
several minibatches."""
import numpy as np
import pytest

from util import small_problem, relerr
from sagecal_b200 import synth
from sagecal_b200.dirac_api import make_barr

pytestmark = pytest.mark.gpu


def multichannel(pr, freqs, seed=3, outliers=0.02):
    """coh [chan][row][M][4] and data [chan][row][8] of a problem at several frequencies, one set of
    true Jones for all channels"""
    rng = np.random.default_rng(seed)
    cohs, xs = [], []
    for f in freqs:
        coh = synth.coherencies(pr.u, pr.v, pr.w, pr.clusters, f, pr.fdelta)
        x = synth.apply_jones(coh, pr.jones_true, pr.sta1, pr.sta2, pr.N, pr.nchunk, pr.flag)
        sig = 1e-2 * np.median(np.abs(x[x != 0]))
        x = x + rng.normal(0, sig, x.shape)
        bad = rng.uniform(0, 1, x.shape) < outliers
        x[bad] += rng.normal(0, 20 * sig, int(bad.sum()))
        x.reshape(-1, 8)[pr.flag != 0] = 0.0
        cohs.append(coh)
        xs.append(x)
    return np.concatenate(cohs), np.concatenate(xs)


@pytest.mark.parametrize("consensus", [False, True], ids=["visibilities", "consensus"])
def test_minibatch_lbfgs_matches_reference(api, ref, consensus):
    b = small_problem(N=9, M=3, tilesz=8, seed=97, kmean=1.0, nchunk=[1, 2, 1])
    pr = b.pr
    freqs = np.array([146e6, 150e6, 154e6])
    nmb, nepoch = 2, 3
    T = pr.tilesz // nmb
    batches = []
    for mb in range(nmb):
        rows = slice(mb * T * pr.Nbase, (mb + 1) * T * pr.Nbase)
        sub = synth.Problem.__new__(synth.Problem)
        sub.__dict__.update(pr.__dict__)
        sub.u, sub.v, sub.w = pr.u[rows], pr.v[rows], pr.w[rows]
        sub.sta1, sub.sta2, sub.flag = pr.sta1[rows], pr.sta2[rows], pr.flag[rows]
        coh, x = multichannel(sub, freqs, seed=5 + mb)
        batches.append((sub, coh, x))
    m = b.m
    rng = np.random.default_rng(8)
    Y = 0.1 * rng.normal(0, 1, m) if consensus else None
    Z = pr.jones_true + 0.05 * rng.normal(0, 1, m) if consensus else None
    rho = rng.uniform(1.0, 10.0, pr.Mt) if consensus else None
    out = []
    for lib in (ref, api):
        pt = lib.persist_init(nmb, m, 8 * T * pr.Nbase * len(freqs), 5)
        pp = pr.pp0.copy()
        hist = []
        for ep in range(nepoch):
            for mb, (sub, coh, x) in enumerate(batches):
                barr = make_barr(sub.sta1, sub.sta2, sub.flag)
                r = lib.bfgsfit_minibatch(sub.u, sub.v, sub.w, x.copy(), pr.N, pr.Nbase, T, barr,
                                          b.sky, coh, pp, freqs, pt, max_lbfgs=3, lbfgs_m=5,
                                          robust_nu=5.0, nmb=mb, totalmb=nmb, Y=Y, Z=Z, rho=rho)
                hist.append((r, pp.copy()))
        lib.persist_clear(pt)
        out.append(hist)
    for (rr, ppr), (rg, ppg) in zip(*out):
        assert abs(rr[0] - rg[0]) <= 1e-9 * abs(rr[0])
        assert abs(rr[1] - rg[1]) <= 1e-7 * abs(rr[1])
        assert relerr(ppg, ppr) < 1e-6, relerr(ppg, ppr)
    # it does calibrate: the cost of the last call is well below that of the first
    assert out[1][-1][0][1] < 0.5 * out[1][0][0][0]
