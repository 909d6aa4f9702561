"""GPU parity at the station count of BASELINE.json config 4: 512 stations, 130816 baselines, 8N = 4096
unknowns per cluster (the damped solves leave the 512-row cluster Cholesky), reduced to 2 timeslots
and 2 clusters so that the CPU restatement could produce a golden (tests/golden/n512, generator
committed; the compiled reference would need a 68 GB dense Jacobian).  Also: assembly of J^T J and
J^T e at N=512 against the restatement's O(rows) normal equations."""
import ast
import os
import sys

import numpy as np
import pytest

from util import relerr

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = os.path.join(HERE, "golden", "n512", "lm.npz")


@pytest.fixture(scope="module")
def prob():
    import make_golden_n512 as gen
    return gen, gen.build()


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden not generated")
def test_sagefit_n512_matches_golden(api, prob):
    from sagecal_b200.dirac_api import SkyModel, make_barr
    gen, pr = prob
    g = np.load(GOLD)
    assert np.allclose(gen.fingerprint(pr), g["fingerprint"], rtol=1e-11, atol=0)
    kw = ast.literal_eval(str(g["args"]))
    barr = make_barr(pr.sta1, pr.sta2, pr.flag)
    sky = SkyModel(pr.clusters, pr.N)
    x, pp = pr.x.copy(), pr.pp0.copy()
    out = api.sagefit_visibilities(pr.u, pr.v, pr.w, x, pr.N, pr.Nbase, pr.tilesz, barr, sky, pr.coh,
                                   pp, **kw)
    want = g["out_scalars"]
    assert out[0] == int(want[0])
    assert abs(out[2] - want[2]) <= 1e-10 * want[2]
    assert relerr(pp, g["out_pp"]) < 1e-5, relerr(pp, g["out_pp"])
    assert abs(out[3] - want[3]) <= 1e-5 * want[3]


@pytest.mark.parametrize("name", ["rtr", "rtr4"])
def test_sagefit_n512_rtr_matches_golden(api, prob, name):
    """robust RTR (the reference driver's default solver) and RSD + RTR at 512 stations: the Jones
    travel through device memory (more than 64 stations) and a 16-lane group walks 8 baseline ends"""
    from sagecal_b200.dirac_api import SkyModel, make_barr
    gold = os.path.join(HERE, "golden", "n512", name + ".npz")
    if not os.path.exists(gold):
        pytest.skip("golden not generated")
    gen, pr = prob
    g = np.load(gold)
    assert np.allclose(gen.fingerprint(pr), g["fingerprint"], rtol=1e-11, atol=0)
    kw = ast.literal_eval(str(g["args"]))
    x, pp = pr.x.copy(), pr.pp0.copy()
    out = api.sagefit_visibilities(pr.u, pr.v, pr.w, x, pr.N, pr.Nbase, pr.tilesz,
                                   make_barr(pr.sta1, pr.sta2, pr.flag), SkyModel(pr.clusters, pr.N),
                                   pr.coh, pp, **kw)
    want = g["out_scalars"]
    assert out[0] == int(want[0])
    assert abs(out[1] - want[1]) < 1e-9
    assert relerr(pp, g["out_pp"]) < 1e-5, relerr(pp, g["out_pp"])
    assert abs(out[3] - want[3]) <= 1e-5 * want[3]


def test_normal_equations_n512(api, prob):
    """J^T J (4096 x 4096) and J^T e of one cluster at N=512 against the O(rows) restatement"""
    import orcdirac
    from sagecal_b200 import lib as blib
    from sagecal_b200.dirac_api import SkyModel, make_barr
    if not orcdirac.available():
        pytest.skip("oracle/liboracle.so not built")
    gen, pr = prob
    barr = make_barr(pr.sta1, pr.sta2, pr.flag)
    sky = SkyModel(pr.clusters, pr.N)
    rng = np.random.default_rng(3)
    n8 = 8 * pr.N
    pblk = pr.pp0[:n8] + 0.1 * rng.normal(0, 1, n8)
    orc = orcdirac.Oracle(pr)
    c_o, JTJ_o, JTe_o = orc.normal_eq(1, 0, pr.tilesz, pblk, pr.x)
    with blib.DeviceProblem(api, pr.N, pr.Nbase, pr.tilesz, barr, sky, pr.coh, pr.x) as dp:
        c_g, JTJ_g, JTe_g = dp.normal_eq(1, 0, pblk, pr.x)
    assert abs(c_g - c_o) <= 1e-12 * c_o
    assert relerr(JTe_g, JTe_o) < 1e-11
    assert relerr(JTJ_g, JTJ_o) < 1e-11
