import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import ctypes as C
    import numpy as np
    from sagecal_b200 import lib as blib, synth
    from sagecal_b200.dirac_api import SkyModel, make_barr
    api = blib.load()
    N, M, T = (int(v) for v in sys.argv[2].split(","))
    pr = synth.make_problem(N=N, M=M, tilesz=T, radius=40e3, seed=5, kmean=2.0, with_data=False)
    pr.x = np.zeros(8 * pr.Nbase1)
    barr = make_barr(pr.sta1, pr.sta2, pr.flag); sky = SkyModel(pr.clusters, pr.N)
    dp = blib.DeviceProblem(api, pr.N, pr.Nbase, pr.tilesz, barr, sky, None, pr.x)
    dp.precalculate(pr.u, pr.v, pr.w, pr.freq0, pr.fdelta)
    pp = pr.jones_true.copy()
    dp.cost(pp)  # uploads Jones
    api.lib.dirac_b200_bench_line_setup.restype = C.c_double
    api.lib.dirac_b200_bench_line_setup.argtypes = [C.c_void_p, C.c_int]
    api.lib.dirac_b200_bench_grad.restype = C.c_double
    api.lib.dirac_b200_bench_grad.argtypes = [C.c_void_p, C.c_int]
    us = api.lib.dirac_b200_bench_line_setup(dp.h, 5)
    by = pr.Nbase1 * (64.0 * M + 65 + 192)
    ug = api.lib.dirac_b200_bench_grad(dp.h, 5)
    print(json.dumps({"shape": sys.argv[2], "cfg": os.environ.get("DIRAC_B200_SA_CFG1", "0"), "us": us, "GBps": by / us / 1e3,
                      "gcfg": os.environ.get("DIRAC_B200_GRADS_CFG", "0"), "grad_us": ug, "grad_GBps": pr.Nbase1 * (64.0 * M + 65) / ug / 1e3}))
    sys.exit(0)
for shape in ("62,64,120", "512,32,120"):
    for cfg in (0, 1, 2, 3, 4, 5):
        env = dict(os.environ); env["DIRAC_B200_GRADS_CFG"] = str(cfg)
        out = subprocess.run([sys.executable, __file__, "child", shape], env=env, capture_output=True, text=True)
        print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
