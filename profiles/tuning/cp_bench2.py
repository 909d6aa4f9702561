import sys, os; sys.path.insert(0,'/root/repo')
import ctypes as C, numpy as np
from sagecal_b200 import synth, lib as blib
from sagecal_b200.dirac_api import SkyModel, make_barr
api = blib.load(); L = api.lib
L.dirac_b200_bench_cluster_pass.restype = C.c_double
L.dirac_b200_bench_cluster_pass.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
L.dirac_b200_bench_grad.restype = C.c_double
L.dirac_b200_bench_grad.argtypes = [C.c_void_p, C.c_int]
pr = synth.make_problem(N=62, M=64, tilesz=120, radius=40e3, seed=5, kmean=2.0)
barr = make_barr(pr.sta1, pr.sta2, pr.flag); sky = SkyModel(pr.clusters, pr.N)
dp = blib.DeviceProblem(api, pr.N, pr.Nbase, pr.tilesz, barr, sky, pr.coh, pr.x)
dp.cost(pr.pp0)
for mode in (0, 1, 2, 3):
    L.dirac_b200_bench_grad(dp.h, -1 - mode)
    us = L.dirac_b200_bench_cluster_pass(dp.h, 0, 1, 1, 0, 10, 300)
    print('dbg', mode, 'TRIAL+grad tslice 10', round(us, 2), 'us')
