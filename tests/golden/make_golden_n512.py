"""Golden outputs at the STATION COUNT of BASELINE.json config 4 (512 stations, 130816 baselines,
8N = 4096 unknowns per cluster) with a reduced interval (2 timeslots, 2 clusters), from the CPU
restatement oracle/liboracle.so (pinned to the compiled reference by tests/test_oracle_vs_ref.py and
tests/test_oracle_c2r.py; the compiled reference itself would need a 68 GB dense Jacobian here).
Stored: solved Jones, scalars, input fingerprint; inputs are regenerated from the seed.

    python tests/golden/make_golden_n512.py [lm] [rtr] [rtr4]   (a few minutes: four 4096^3/3 factorisations
                                                        per sweep in plain C)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "n512")
SHAPE = dict(N=512, M=2, tilesz=2, radius=75e3, seed=20260921 + 4, kmean=2.0)
SOLVE = dict(max_emiter=2, max_iter=2, max_lbfgs=4, lbfgs_m=5, linsolv=0, solver_mode=1, nulow=2.0,
             nuhigh=30.0)


def fingerprint(pr):
    return np.array([np.sum(pr.x), np.sum(np.abs(pr.x)), np.sum(pr.coh.real), np.sum(pr.coh.imag),
                     np.sum(np.abs(pr.coh)), float(np.sum(pr.flag)), np.sum(pr.u), np.sum(pr.w)])


def build():
    from sagecal_b200 import synth
    return synth.make_problem(**SHAPE)


#: the same problem under the reference driver's default solver (robust RTR, solver_mode 5) and under
#: RSD + RTR (4): no 4096 x 4096 systems at all on this path
VARIANTS = {"lm": SOLVE, "rtr": dict(SOLVE, solver_mode=5), "rtr4": dict(SOLVE, solver_mode=4)}


def main():
    import orcdirac
    os.makedirs(OUT, exist_ok=True)
    pr = build()
    o = orcdirac.Oracle(pr)
    for name in (sys.argv[1:] or ["lm"]):
        solve = VARIANTS[name]
        x, pp = pr.x.copy(), pr.pp0.copy()
        t0 = time.time()
        out = o.sagefit(x, pp, **solve)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), args=np.array(repr(solve)), out_pp=pp,
                            out_scalars=np.array(out, dtype=np.float64),
                            fingerprint=fingerprint(pr))
        print(name, out, "%.1f s" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
