"""driver for the `ncu --set full` capture of the RTR-family kernels: the C3 shape (62 stations,
120 timeslots) with 4 clusters, one SAGE sweep under solver_mode 5"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sagecal_b200 import synth, lib as blib
from sagecal_b200.dirac_api import SkyModel, make_barr
api = blib.load()
pr = synth.make_problem(N=62, M=4, tilesz=120, radius=40e3, seed=20260924, kmean=2.0, outliers=0.02)
barr = make_barr(pr.sta1, pr.sta2, pr.flag)
sky = SkyModel(pr.clusters, pr.N)
dp = blib.DeviceProblem(api, pr.N, pr.Nbase, pr.tilesz, barr, sky, pr.coh, pr.x)
pp = pr.pp0.copy()
print(dp.sagefit(pp, None, max_emiter=1, max_iter=2, max_lbfgs=0, lbfgs_m=7, solver_mode=5))
