"""Golden outputs at the FULL benchmark shapes (BASELINE.json configs 2 and 3: 62 stations, 1891
baselines, 64 clusters, 120 timeslots) from the CPU restatement `oracle/liboracle.so` — the compiled
reference cannot run these shapes (7.2 GB dense Jacobian and ~0.9 PFLOP of dgemm per cluster
iteration), the O(rows) restatement can (about 10-20 minutes, single thread).  The restatement is
pinned to the compiled reference on small shapes by tests/test_oracle_vs_ref.py and on the reduced
C2/C3 shape by tests/test_oracle_c2r.py.  Stored: solved Jones, scalars, input fingerprint; the
inputs are regenerated from the seed (sagecal_b200.synth.make_config).  bench.py compares the
solution of its first warm-up step with these (the `parity` object of the bench line).

    python tests/golden/make_golden_full.py C2 | C3 | C3os | C2rtr | C3rtr | C3nsd
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "full")
SOLVE = dict(max_emiter=3, max_iter=2, max_lbfgs=10, lbfgs_m=7, linsolv=0, nulow=2.0, nuhigh=30.0)
CASES = {"C2": ("C2", 1), "C3": ("C3", 2), "C3os": ("C3", 3),
         # RTR family (oracle/rtr_harness.cpp: rtr_algo.h on the per-row evaluators)
         "C2rtr": ("C2", 4), "C3rtr": ("C3", 5), "C3nsd": ("C3", 6),
         # the SAGE stage alone (no LBFGS): the Gaussian LBFGS stage differentiates the cost
         # numerically with a step of 1e-9..1e-6 (lbfgs.c:546), so its iterates carry the rounding of
         # the cost sum; the stage before it is reproducible to 1e-12
         "C2lm": ("C2", 1, dict(max_lbfgs=0))}


def fingerprint(pr):
    return np.array([np.sum(pr.x), np.sum(np.abs(pr.x)), np.sum(pr.coh.real), np.sum(pr.coh.imag),
                     np.sum(np.abs(pr.coh)), float(np.sum(pr.flag)), np.sum(pr.u), np.sum(pr.w)])


def main():
    import orcdirac
    from sagecal_b200 import synth
    os.makedirs(OUT, exist_ok=True)
    for name in sys.argv[1:]:
        cfg, mode = CASES[name][:2]
        over = CASES[name][2] if len(CASES[name]) > 2 else {}
        pr = synth.make_config(cfg)
        o = orcdirac.Oracle(pr)
        x, pp = pr.x.copy(), pr.pp0.copy()
        t0 = time.time()
        kw = dict(SOLVE)
        kw.update(over)
        kw["solver_mode"] = mode
        out = o.sagefit(x, pp, **kw)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), config=np.array(cfg),
                            args=np.array(repr(kw)), out_pp=pp,
                            out_scalars=np.array(out, dtype=np.float64),
                            out_x_fp=np.array([np.sum(x), np.sum(np.abs(x)), np.max(np.abs(x))]),
                            fingerprint=fingerprint(pr))
        print(name, out, "%.1f s" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
