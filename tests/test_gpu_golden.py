"""GPU: the product against the committed golden vectors (reference outputs) through the C-ABI."""
import numpy as np
import pytest

import golden_util
from util import relerr
from sagecal_b200 import lib as blib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_util.names())
def test_product_matches_golden(api, name):
    b, args, g = golden_util.load(name)
    pr = b.pr
    pp = g["pp_probe"]
    with blib.DeviceProblem(api, pr.N, pr.Nbase, pr.tilesz, b.barr, b.sky, pr.coh, pr.x) as dp:
        _, model = dp.predict(pp, out_mode=2)
        assert relerr(model, g["model_probe"]) < 1e-13
        assert abs(dp.cost(pp) - float(g["cost_gauss"])) <= 1e-12 * float(g["cost_gauss"])
        assert abs(dp.cost(pp, True, 3.0) - float(g["cost_robust"])) <= 1e-12 * float(g["cost_robust"])
        assert relerr(dp.grad(pp), g["grad_gauss"]) < 1e-11
        assert relerr(dp.grad(pp, True, 3.0), g["grad_robust"]) < 1e-11
    x, p = pr.x.copy(), pr.pp0.copy()
    rv, nu, r0, r1 = api.sagefit_visibilities(pr.u, pr.v, pr.w, x, pr.N, pr.Nbase, pr.tilesz,
                                              b.fresh_barr(), b.sky, pr.coh, p, **args)
    want = g["out_scalars"]
    assert rv == int(want[0]) and abs(nu - want[1]) < 1e-9
    assert abs(r0 - want[2]) <= 1e-10 * want[2] and abs(r1 - want[3]) <= 1e-5 * want[3]
    assert relerr(p, g["out_pp"]) < 1e-5          # north_star: 1e-5 relative on the solved Jones
    assert relerr(x, g["out_x"]) < 1e-5 * np.max(np.abs(pr.x)) / np.max(np.abs(g["out_x"]))
