"""Golden vectors at the "reduced C2/C3" shape BASELINE.md 5.5 names (62 stations, 1891 baselines,
8 clusters, 10 timeslots): the compiled reference (oracle/_ref) solves seeded synthetic problems with
the benchmark's solver settings; only the OUTPUTS (solved Jones, scalars) and a fingerprint of the
inputs are stored — the inputs are regenerated from the seed by sagecal_b200.synth (9.7 MB of
coherencies per case would not belong in git).  Takes minutes per case (dense Jacobian + dgemm).

    python tests/golden/make_golden_c2r.py [case ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c2r")

SHAPE = dict(N=62, M=8, tilesz=10, radius=40e3, kmean=2.0)
SOLVE = dict(max_emiter=3, max_iter=2, max_lbfgs=10, lbfgs_m=7, linsolv=0, nulow=2.0, nuhigh=30.0,
             randomize=0)
CASES = {
    # name: (problem overrides, entry point, solver arguments)
    "lm": (dict(seed=20260921 + 2), "sagefit_visibilities", dict(solver_mode=1)),
    "oslm": (dict(seed=20260921 + 12), "sagefit_visibilities", dict(solver_mode=0)),
    "rlm": (dict(seed=20260921 + 3, outliers=0.02), "sagefit_visibilities", dict(solver_mode=2)),
    # OS robust LM rejects trial steps down to rounding level (dF ~ 1e-11 on a cost of ~3e4), where
    # the accept/reject decision is rounding noise that nevertheless resets mu and nu: the compiled
    # reference and the CPU restatement part ways on about half of the seeds tried (Jones differing
    # by 1e-3 .. 1e-1); this seed is one on which they agree to 4e-12.  The GPU test holds the Jones
    # to 1e-5 only if it took the same branches (see test_reduced_c2_matches_reference_golden).
    "osrlm": (dict(seed=101, outliers=0.02), "sagefit_visibilities", dict(solver_mode=3)),
    "hybrid": (dict(seed=20260921 + 14, nchunk=[1, 2, 1, 1, 3, 1, 1, 1]), "sagefit_visibilities",
               dict(solver_mode=1)),
    "bfgs_robust": (dict(seed=20260921 + 15, outliers=0.02), "bfgsfit_visibilities",
                    dict(solver_mode=2, mean_nu=5.0, max_lbfgs=10, lbfgs_m=7)),
}


def fingerprint(pr):
    """sums that pin the regenerated inputs (compared to 1e-12 relative by the tests)"""
    return np.array([np.sum(pr.x), np.sum(np.abs(pr.x)), np.sum(pr.coh.real), np.sum(pr.coh.imag),
                     np.sum(np.abs(pr.coh)), float(np.sum(pr.flag)), np.sum(pr.u), np.sum(pr.w)])


def build(name):
    from sagecal_b200 import synth
    from util import Bound
    over, fn, args = CASES[name]
    shape = dict(SHAPE)
    shape.update(over)
    start = shape.pop("start", "identity")
    b = Bound(synth.make_problem(**shape))
    if start == "near":
        rng = np.random.default_rng(shape["seed"] + 1000)
        b.pr.pp0 = b.pr.jones_true + 0.03 * rng.normal(0, 1, b.pr.jones_true.shape)
    return b, fn, args


def main():
    import refdirac
    ref = refdirac.load()
    os.makedirs(OUT, exist_ok=True)
    for name in (sys.argv[1:] or list(CASES)):
        b, fn, args = build(name)
        pr = b.pr
        x, pp = pr.x.copy(), pr.pp0.copy()
        t0 = time.time()
        if fn == "sagefit_visibilities":
            kw = dict(SOLVE)
            kw.update(args)
            out = ref.sagefit_visibilities(pr.u, pr.v, pr.w, x, pr.N, pr.Nbase, pr.tilesz,
                                           b.fresh_barr(), b.sky, pr.coh, pp, Nt=8, **kw)
        else:
            kw = dict(args)
            out = ref.bfgsfit_visibilities(pr.u, pr.v, pr.w, x, pr.N, pr.Nbase, pr.tilesz,
                                           b.fresh_barr(), b.sky, pr.coh, pp, Nt=8, **kw)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), args=np.array(repr(kw)),
                            fn=np.array(fn), out_pp=pp, out_scalars=np.array(out, dtype=np.float64),
                            out_x_fp=np.array([np.sum(x), np.sum(np.abs(x)), np.max(np.abs(x))]),
                            fingerprint=fingerprint(pr))
        print(name, out, "%.1f s" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
