"""CPU: the restated oracle against the committed golden vectors (reference outputs)."""
import numpy as np
import pytest

import golden_util
import orcdirac
from util import relerr


@pytest.mark.parametrize("name", golden_util.names())
def test_oracle_matches_golden(name):
    if not orcdirac.available():
        pytest.skip("oracle/liboracle.so not built")
    b, args, g = golden_util.load(name)
    pr = b.pr
    orc = orcdirac.Oracle(pr)
    pp = g["pp_probe"]
    assert relerr(orc.predict_full(pp), g["model_probe"]) < 1e-14
    assert abs(orc.cost(pp, pr.x, False) - float(g["cost_gauss"])) <= 1e-12 * float(g["cost_gauss"])
    assert abs(orc.cost(pp, pr.x, True, 3.0) - float(g["cost_robust"])) <= 1e-12 * float(g["cost_robust"])
    assert relerr(orc.grad(pp, pr.x, False), g["grad_gauss"]) < 1e-12
    assert relerr(orc.grad(pp, pr.x, True, 3.0), g["grad_robust"]) < 1e-12
    x, p = pr.x.copy(), pr.pp0.copy()
    rv, nu, r0, r1 = orc.sagefit(x, p, **args)
    want = g["out_scalars"]
    assert rv == int(want[0]) and abs(nu - want[1]) < 1e-9
    assert abs(r0 - want[2]) <= 1e-12 * want[2] and abs(r1 - want[3]) <= 1e-6 * want[3]
    assert relerr(p, g["out_pp"]) < 1e-6
