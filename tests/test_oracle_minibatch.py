"""CPU: the minibatch LBFGS control flow the product runs on the host
(sagecal_b200/csrc/minibatch_algo.h) on top of the oracle's per-row Student's-t cost and gradient
(oracle/minibatch_harness.cpp), against the compiled reference's bfgsfit_minibatch_visibilities /
bfgsfit_minibatch_consensus over epochs of minibatches with persistent state.  Runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

import orcdirac
from util import small_problem, relerr
from sagecal_b200 import synth
from sagecal_b200.dirac_api import make_barr, dptr
from test_gpu_minibatch import multichannel

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "oracle", "libminibatch_harness.so")


@pytest.mark.parametrize("consensus", [False, True], ids=["visibilities", "consensus"])
def test_minibatch_control_flow_matches_reference(ref, consensus):
    if not os.path.exists(LIB) or not orcdirac.available():
        pytest.skip("oracle/libminibatch_harness.so not built (make -C oracle)")
    b = small_problem(N=8, M=3, tilesz=8, seed=98, kmean=1.0, nchunk=[1, 2, 1])
    pr = b.pr
    freqs = np.array([147e6, 153e6])
    nmb, nepoch = 2, 3
    T = pr.tilesz // nmb
    batches = []
    for mb in range(nmb):
        rows = slice(mb * T * pr.Nbase, (mb + 1) * T * pr.Nbase)
        sub = synth.Problem.__new__(synth.Problem)
        sub.__dict__.update(pr.__dict__)
        sub.u, sub.v, sub.w = pr.u[rows], pr.v[rows], pr.w[rows]
        sub.sta1, sub.sta2, sub.flag = pr.sta1[rows], pr.sta2[rows], pr.flag[rows]
        sub.tilesz = T
        coh, x = multichannel(sub, freqs, seed=15 + mb)
        batches.append((sub, coh, x))
    m = b.m
    rng = np.random.default_rng(9)
    Y = 0.1 * rng.normal(0, 1, m) if consensus else None
    Z = pr.jones_true + 0.05 * rng.normal(0, 1, m) if consensus else None
    rho = rng.uniform(1.0, 10.0, pr.Mt) if consensus else None
    # reference
    pt = ref.persist_init(nmb, m, 8 * T * pr.Nbase * len(freqs), 5)
    ppr = pr.pp0.copy()
    hist_ref = []
    for ep in range(nepoch):
        for mb, (sub, coh, x) in enumerate(batches):
            barr = make_barr(sub.sta1, sub.sta2, sub.flag)
            r = ref.bfgsfit_minibatch(sub.u, sub.v, sub.w, x.copy(), pr.N, pr.Nbase, T, barr, b.sky,
                                      coh, ppr, freqs, pt, max_lbfgs=3, lbfgs_m=5, robust_nu=5.0,
                                      nmb=mb, totalmb=nmb, Y=Y, Z=Z, rho=rho)
            hist_ref.append((r, ppr.copy()))
    ref.persist_clear(pt)
    # the product's control flow on the oracle's evaluators
    H = C.CDLL(LIB)
    H.harness_persist_new.restype = C.c_void_p
    H.harness_persist_new.argtypes = [C.c_int, C.c_int]
    H.harness_persist_free.argtypes = [C.c_void_p]
    vpp = C.POINTER(C.c_void_p)
    dpp = C.POINTER(C.POINTER(C.c_double))
    dp = C.POINTER(C.c_double)
    H.harness_minibatch_fit.argtypes = [vpp, dpp, C.c_int, C.c_int, C.c_int, C.c_long, dp, dp, dp, dp,
                                        C.c_int, C.c_int, C.c_double, dp, dp, C.c_void_p]
    pth = H.harness_persist_new(m, 5)
    ppo = pr.pp0.copy()
    n1 = 4 * pr.M * T * pr.Nbase       # complex coherencies per channel
    for ep in range(nepoch):
        for mb, (sub, coh, x) in enumerate(batches):
            orcs, xs = [], []
            for c in range(len(freqs)):
                orcs.append(orcdirac.Oracle(sub, coh=coh[c * n1:(c + 1) * n1]))
                xs.append(np.ascontiguousarray(x[c * 8 * T * pr.Nbase:(c + 1) * 8 * T * pr.Nbase]))
            Parr = (C.c_void_p * len(freqs))(*[C.cast(C.pointer(o.P), C.c_void_p) for o in orcs])
            Xarr = (dp * len(freqs))(*[dptr(a) for a in xs])
            r0, r1 = C.c_double(0), C.c_double(0)
            H.harness_minibatch_fit(Parr, Xarr, len(freqs), pr.N, pr.Mt, T * pr.Nbase, dptr(ppo),
                                    dptr(Y) if consensus else None, dptr(Z) if consensus else None,
                                    dptr(rho) if consensus else None, 3, 5, 5.0, C.byref(r0),
                                    C.byref(r1), pth)
            (rr, ppw) = hist_ref.pop(0)
            assert abs(r0.value - rr[0]) <= 1e-10 * abs(rr[0])
            assert abs(r1.value - rr[1]) <= 1e-8 * abs(rr[1])
            assert relerr(ppo, ppw) < 1e-8, (ep, mb, relerr(ppo, ppw))
    H.harness_persist_free(pth)
