"""Where the end-to-end C4 step goes beyond the resident solve: wall time (host clock around
device-synchronised stages) of create / precalculate / sagefit (+ D2H) / destroy on one GPU.
    python profiles/tuning/e2e_stages.py [clusters_per_gpu]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from sagecal_b200 import lib as blib, synth
    from sagecal_b200.dirac_api import SkyModel, make_barr
    import bench
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    api = blib.load()
    shape = dict(bench.workload_shape("C4"))
    shape["M"] = M
    pr = synth.make_problem(with_data=False, **shape)
    barr = make_barr(pr.sta1, pr.sta2, pr.flag)
    sky = SkyModel(pr.clusters, pr.N)
    x = torch.from_numpy(np.zeros(8 * pr.Nbase1)).pin_memory().numpy()
    pp = torch.from_numpy(pr.pp0.copy()).pin_memory().numpy()
    xo = torch.from_numpy(np.zeros(8 * pr.Nbase1)).pin_memory().numpy()
    solve = dict(bench.SOLVE)
    out = []
    for it in range(3):
        t = [time.perf_counter()]

        def mark():
            torch.cuda.synchronize()
            t.append(time.perf_counter())
        dp = blib.DeviceProblem(api, pr.N, pr.Nbase, pr.tilesz, barr, sky, None, x)
        mark()
        dp.precalculate(pr.u, pr.v, pr.w, pr.freq0, pr.fdelta)
        mark()
        if it == 0:  # data = model + nothing: only the timings matter here
            api.lib.dirac_b200_predict(dp.h, blib.dptr(pr.jones_true), blib.dptr(x), 2, 0, 0.0)
            dp.set_data(x)
            torch.cuda.synchronize()
            t[-1] = time.perf_counter()
        pp[:] = pr.pp0
        dp.sagefit(pp, xo, **solve)
        mark()
        dp.close()
        mark()
        d = np.diff(t) * 1e3
        out.append(dict(create_ms=d[0], precalculate_ms=d[1], sagefit_d2h_ms=d[2], destroy_ms=d[3]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
