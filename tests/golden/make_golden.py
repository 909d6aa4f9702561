"""Generates the golden vectors under tests/golden/ by running the compiled reference
(oracle/_ref/libdirac_ref.so, built from /root/reference by oracle/Makefile) on seeded synthetic
inputs.  The reference ships no known-answer tests for this path, so these vectors
are what travels to the GPU box.  Run here (needs /root/reference to have been built once):

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refdirac  # noqa: E402
from util import small_problem, perturbed_jones  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "c1_lm": (dict(N=8, M=2, tilesz=10, seed=20260922), dict(solver_mode=1, max_iter=5)),
    "oslm": (dict(N=8, M=3, tilesz=20, seed=33, kmean=1.0), dict(solver_mode=0, max_iter=4)),
    "rlm": (dict(N=8, M=2, tilesz=10, seed=34, outliers=0.02), dict(solver_mode=2, max_iter=3)),
    "osrlm": (dict(N=8, M=2, tilesz=20, seed=35, outliers=0.02), dict(solver_mode=3, max_iter=3)),
    "hybrid": (dict(N=8, M=3, tilesz=10, seed=36, nchunk=[1, 2, 5]), dict(solver_mode=1, max_iter=3)),
}


def main():
    ref = refdirac.load()
    for name, (prob, args) in CASES.items():
        b = small_problem(**prob)
        pr = b.pr
        kw = dict(max_emiter=3, max_lbfgs=6, lbfgs_m=5, randomize=0)
        kw.update(args)
        x, pp = pr.x.copy(), pr.pp0.copy()
        rv, nu, r0, r1 = ref.sagefit_visibilities(pr.u, pr.v, pr.w, x, pr.N, pr.Nbase, pr.tilesz,
                                                  b.fresh_barr(), b.sky, pr.coh, pp, **kw)
        ppj = perturbed_jones(pr)
        md = ref.me_data(pr.N, pr.Nbase, pr.tilesz, b.barr, b.sky, pr.coh, robust_nu=3.0)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            prob=np.array(repr(prob)), args=np.array(repr(kw)),
            u=pr.u, v=pr.v, w=pr.w, x=pr.x, coh=pr.coh, flag=pr.flag, nchunk=np.array(pr.nchunk),
            pp0=pr.pp0,
            out_pp=pp, out_x=x, out_scalars=np.array([rv, nu, r0, r1]),
            pp_probe=ppj, model_probe=ref.predict_full(ppj, md, b.n),
            cost_gauss=ref.cost(ppj, pr.x, md), cost_robust=ref.cost(ppj, pr.x, md, robust=True),
            grad_gauss=ref.grad(ppj, pr.x, md), grad_robust=ref.grad(ppj, pr.x, md, robust=True))
        print(name, rv, nu, r0, r1)


if __name__ == "__main__":
    main()
