"""GPU parity on the code paths the BENCHMARK shape runs but small problems do not reach by
themselves:
  * k_cluster_pass_lin with several rows per CTA: the 5-stage TMA ring is refilled, the mbarrier
    parity flips, flag bits of rows >= 5 are used (at C2 a CTA owns 7 rows; the launcher gives every
    small problem 1 row per CTA, so the tests force the slicing with the `cp_rows` option);
  * k_stream_all (full predict, LBFGS line model) with more clusters per warp than ring stages
    (M >= 9 refills a stage; C2 has 22 clusters per warp);
  * hybrid clusters whose chunks do not tile the interval evenly (row-based chunk map of the hidden
    data);
  * the reduced C2/C3 shape of BASELINE.md 5.5 (62 stations, 8 clusters, 10 timeslots) against
    golden outputs of the compiled reference (tests/golden/c2r, generator committed).
All against the compiled reference (oracle/_ref) or its golden outputs, Jones within 1e-5."""
import ast
import os

import numpy as np
import pytest

from util import small_problem, relerr
from test_gpu_solvers import run_both, JONES_TOL

pytestmark = pytest.mark.gpu


@pytest.fixture
def cp_rows(api):
    def setter(v):
        api.set_option("cp_rows", v)
    yield setter
    api.set_option("cp_rows", 0)


RING_CASES = [
    # 20 timeslots, 7 rows per CTA: slices of 7, 7, 6 rows -> stages 0 and 1 are refilled (parity 1)
    ("lm-7rows", 7, dict(N=12, M=3, tilesz=20, seed=71, kmean=1.0, flag_frac=0.2),
     dict(solver_mode=1, max_iter=3)),
    # one CTA walks all 24 rows: the ring wraps four times, flag bits up to bit 23
    ("lm-24rows", 24, dict(N=9, M=2, tilesz=24, seed=72, flag_frac=0.3, uvcut_frac=0.02),
     dict(solver_mode=1, max_iter=3)),
    # more than 256 baselines: several baseline groups, the last one ragged (N=30: 435 = 256 + 179)
    ("lm-2groups", 6, dict(N=30, M=2, tilesz=12, seed=73, kmean=1.0, flag_frac=0.1),
     dict(solver_mode=1, max_iter=2)),
    # hybrid chunks (evenly tiling) with several rows per CTA
    ("lm-hybrid-rows", 4, dict(N=12, M=3, tilesz=20, seed=74, nchunk=[2, 1, 5], flag_frac=0.1),
     dict(solver_mode=1, max_iter=3)),
    # OS-LM (subset passes are short: 2 rows each) followed by plain LM on 20 rows
    ("oslm-rows", 6, dict(N=10, M=3, tilesz=20, seed=75, kmean=1.0), dict(solver_mode=0, max_iter=4)),
]


@pytest.mark.parametrize("name,rows,prob,args", RING_CASES, ids=[c[0] for c in RING_CASES])
def test_cluster_pass_ring_matches_reference(api, ref, cp_rows, name, rows, prob, args):
    b = small_problem(**prob)
    kw = dict(max_emiter=3, max_lbfgs=6, lbfgs_m=5, randomize=0)
    kw.update(args)
    cp_rows(rows)
    (rr, xr, ppr), (rg, xg, ppg) = run_both(api, ref, b, **kw)
    assert rr[0] == rg[0]
    assert abs(rr[2] - rg[2]) <= 1e-10 * rr[2]
    assert relerr(ppg, ppr) < JONES_TOL, (name, relerr(ppg, ppr))
    assert abs(rr[3] - rg[3]) <= 1e-5 * rr[3]


def test_cluster_pass_rows_do_not_change_the_answer(api, cp_rows):
    """same problem, 1 / 5 / 6 / 32 rows per CTA: identical up to summation order"""
    b = small_problem(N=14, M=2, tilesz=40, seed=76, flag_frac=0.15)
    pr = b.pr
    sols = []
    for rows in (0, 5, 6, 32):
        cp_rows(rows)
        x, pp = pr.x.copy(), pr.pp0.copy()
        r = api.sagefit_visibilities(pr.u, pr.v, pr.w, x, pr.N, pr.Nbase, pr.tilesz, b.fresh_barr(),
                                     b.sky, pr.coh, pp, max_emiter=2, max_iter=3, max_lbfgs=0,
                                     solver_mode=1)
        sols.append((r, pp))
    for r, pp in sols[1:]:
        assert relerr(pp, sols[0][1]) < 1e-9
        assert abs(r[3] - sols[0][0][3]) <= 1e-9 * sols[0][0][3]


MANY_CLUSTER_CASES = [
    # 20 clusters -> 7 per warp of k_stream_all with 2 stages: every stage refilled three times
    ("m20-lm", dict(N=8, M=20, tilesz=6, seed=81, kmean=1.0), dict(solver_mode=1, max_iter=2)),
    # hybrid clusters among them (Jones change inside a CTA's rows)
    ("m11-hybrid", dict(N=9, M=11, tilesz=8, seed=82, nchunk=[1, 2, 1, 1, 4, 1, 1, 1, 2, 1, 1]),
     dict(solver_mode=1, max_iter=2)),
    ("m13-robust", dict(N=8, M=13, tilesz=10, seed=83, outliers=0.02), dict(solver_mode=2, max_iter=2)),
]


@pytest.mark.parametrize("name,prob,args", MANY_CLUSTER_CASES, ids=[c[0] for c in MANY_CLUSTER_CASES])
def test_many_clusters_match_reference(api, ref, name, prob, args):
    b = small_problem(**prob)
    kw = dict(max_emiter=2, max_lbfgs=8, lbfgs_m=5, randomize=0)
    kw.update(args)
    (rr, xr, ppr), (rg, xg, ppg) = run_both(api, ref, b, **kw)
    assert rr[0] == rg[0]
    assert abs(rr[2] - rg[2]) <= 1e-10 * rr[2]
    assert relerr(ppg, ppr) < JONES_TOL, (name, relerr(ppg, ppr))
    assert abs(rr[3] - rg[3]) <= 1e-5 * rr[3]


def test_many_clusters_predict_cost_grad(api, ref):
    """full predict, both costs and both gradients at M=20 against the reference callbacks"""
    from sagecal_b200 import lib as blib
    from util import perturbed_jones
    b = small_problem(N=10, M=20, tilesz=7, seed=84, kmean=1.0, nchunk=[1] * 17 + [2, 7, 3])
    pr = b.pr
    pp = perturbed_jones(pr)
    md = ref.me_data(pr.N, pr.Nbase, pr.tilesz, b.barr, b.sky, pr.coh, robust_nu=3.0)
    with blib.DeviceProblem(api, pr.N, pr.Nbase, pr.tilesz, b.barr, b.sky, pr.coh, pr.x) as dp:
        _, model = dp.predict(pp, out_mode=2)
        assert relerr(model, ref.predict_full(pp, md, b.n)) < 1e-13
        c = ref.cost(pp, pr.x, md)
        assert abs(dp.cost(pp) - c) <= 1e-12 * c
        c = ref.cost(pp, pr.x, md, robust=True)
        assert abs(dp.cost(pp, True, 3.0) - c) <= 1e-12 * c
        assert relerr(dp.grad(pp), ref.grad(pp, pr.x, md)) < 1e-11
        assert relerr(dp.grad(pp, True, 3.0), ref.grad(pp, pr.x, md, robust=True)) < 1e-11


HYBRID_UNEVEN = [
    # nchunk does not divide tilesz: hidden data / residual with the row-based chunk map
    ("lm-uneven", dict(N=10, M=3, tilesz=10, seed=91, nchunk=[3, 1, 4]), dict(solver_mode=1, max_iter=3)),
    # (every chunk keeps at least one timeslot: the reference callocs 0 bytes for an empty chunk and exits on
    # allocators that return NULL for that)
    ("lm-uneven2", dict(N=12, M=4, tilesz=7, seed=92, nchunk=[2, 3, 1, 4]), dict(solver_mode=1, max_iter=2)),
    ("rlm-uneven", dict(N=9, M=2, tilesz=10, seed=93, nchunk=[3, 1], outliers=0.02),
     dict(solver_mode=2, max_iter=2)),
]


@pytest.mark.parametrize("name,prob,args", HYBRID_UNEVEN, ids=[c[0] for c in HYBRID_UNEVEN])
def test_hybrid_uneven_chunks_match_reference(api, ref, name, prob, args):
    b = small_problem(**prob)
    kw = dict(max_emiter=3, max_lbfgs=6, lbfgs_m=5, randomize=0)
    kw.update(args)
    (rr, xr, ppr), (rg, xg, ppg) = run_both(api, ref, b, **kw)
    assert rr[0] == rg[0]
    assert abs(rr[2] - rg[2]) <= 1e-10 * rr[2]
    assert relerr(ppg, ppr) < JONES_TOL, (name, relerr(ppg, ppr))
    assert abs(rr[3] - rg[3]) <= 1e-5 * rr[3]


# ---------------------------------------------------------------------------------------------
# reduced C2/C3 shape against golden outputs of the compiled reference
# ---------------------------------------------------------------------------------------------
C2R_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2r")
C2R_NAMES = sorted(f[:-4] for f in os.listdir(C2R_DIR) if f.endswith(".npz")) if os.path.isdir(C2R_DIR) else []


@pytest.mark.parametrize("name", C2R_NAMES)
def test_reduced_c2_matches_reference_golden(api, name):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_c2r as gen
    g = np.load(os.path.join(C2R_DIR, name + ".npz"))
    b, fn, _ = gen.build(name)
    pr = b.pr
    fp = gen.fingerprint(pr)
    assert np.allclose(fp, g["fingerprint"], rtol=1e-11, atol=0), "synthetic inputs differ from the golden run"
    kw = ast.literal_eval(str(g["args"]))
    x, pp = pr.x.copy(), pr.pp0.copy()
    api.noise_decisions(reset=True)
    out = getattr(api, fn)(pr.u, pr.v, pr.w, x, pr.N, pr.Nbase, pr.tilesz, b.fresh_barr(), b.sky,
                           pr.coh, pp, **kw)
    want = g["out_scalars"]
    assert out[0] == int(want[0])
    if fn == "sagefit_visibilities":
        assert abs(out[2] - want[2]) <= 1e-10 * want[2]          # res_0
        if name != "osrlm":
            assert abs(out[1] - want[1]) < 1e-9                  # mean nu
            assert abs(out[3] - want[3]) <= 1e-5 * want[3]       # res_1
    else:
        assert abs(out[1] - want[1]) <= 1e-10 * want[1]
        assert abs(out[2] - want[2]) <= 1e-5 * want[2]
    err = relerr(pp, g["out_pp"])
    if name == "osrlm" and err >= JONES_TOL:
        # OS robust LM: accept/reject decisions at rounding level make the iterates of two correct
        # implementations diverge (the compiled reference and its CPU restatement do, see
        # make_golden_c2r.py); then only the quality of the solution is comparable
        assert api.noise_decisions() > 0
        assert abs(out[3] - want[3]) <= 0.05 * want[3]
        pytest.xfail("rounding-level LM decision took another branch than the reference (Jones differ "
                     "by %.1e, res_1 by %.1e relative)" % (err, abs(out[3] - want[3]) / want[3]))
    assert err < JONES_TOL, (name, err)
    xfp = np.array([np.sum(x), np.sum(np.abs(x)), np.max(np.abs(x))])
    assert np.allclose(xfp[1:], g["out_x_fp"][1:], rtol=1e-5)


OS_MISALIGNED = [
    # tile counts that are not a multiple of the 10 ordered subsets (and not < 10): the reference's own
    # pairing of Jacobian rows with residuals of other tiles (clmfit.c:1313-1413), reproduced
    ("oslm-12", dict(N=9, M=2, tilesz=12, seed=101), dict(solver_mode=0, max_iter=3)),
    ("oslm-15", dict(N=10, M=3, tilesz=15, seed=102, kmean=1.0), dict(solver_mode=0, max_iter=4)),
    ("oslm-25", dict(N=8, M=2, tilesz=25, seed=103, flag_frac=0.1), dict(solver_mode=0, max_iter=3)),
    # hybrid: 30 timeslots in 2 chunks of 15 tiles
    ("oslm-hybrid-15", dict(N=9, M=2, tilesz=30, seed=104, nchunk=[2, 1]), dict(solver_mode=0, max_iter=3)),
    ("osrlm-15", dict(N=8, M=2, tilesz=15, seed=105, outliers=0.02), dict(solver_mode=3, max_iter=3)),
]


@pytest.mark.parametrize("name,prob,args", OS_MISALIGNED, ids=[c[0] for c in OS_MISALIGNED])
def test_os_lm_with_the_reference_subset_pairing(api, ref, name, prob, args):
    b = small_problem(**prob)
    kw = dict(max_emiter=3, max_lbfgs=4, lbfgs_m=5, randomize=0)
    kw.update(args)
    api.noise_decisions(reset=True)
    (rr, xr, ppr), (rg, xg, ppg) = run_both(api, ref, b, **kw)
    assert rr[0] == rg[0]
    assert abs(rr[2] - rg[2]) <= 1e-10 * rr[2]
    err = relerr(ppg, ppr)
    if err >= JONES_TOL and api.noise_decisions() > 0:
        pytest.xfail("rounding-level LM decision took another branch than the reference (%.1e)" % err)
    assert err < JONES_TOL, (name, err)
    assert abs(rr[3] - rg[3]) <= 1e-5 * rr[3]
