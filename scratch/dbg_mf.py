import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, refdirac
from util import *
from sagecal_b200 import synth, lib
from sagecal_b200.dirac_api import SkyModel
ref = refdirac.load(); api = lib.load()
def run(spec, gfrac, freqs, M=3):
    b = small_problem(N=10, M=M, tilesz=5, seed=22, kmean=2.0, gaussian_frac=gfrac)
    pr = b.pr
    if spec:
        for cl in pr.clusters:
            K = len(cl["ll"])
            cl["spec_idx"] = np.where(np.arange(K) % 2 == 0, -0.7, 0.0)
            cl["spec_idx1"] = np.full(K, 0.05); cl["spec_idx2"] = np.full(K, -0.01); cl["f0"] = np.full(K, 140e6)
    sky = SkyModel(pr.clusters, pr.N)
    freqs=np.array(freqs)
    xa = np.zeros(8*pr.Nbase1*len(freqs)); xb = xa.copy()
    ref.predict_visibilities_multifreq(pr.u, pr.v, pr.w, xa, pr.N, pr.Nbase, pr.tilesz, b.barr, sky, freqs, pr.fdelta*3, add_to_data=0)
    api.predict_visibilities_multifreq(pr.u, pr.v, pr.w, xb, pr.N, pr.Nbase, pr.tilesz, b.barr, sky, freqs, pr.fdelta*3, add_to_data=0)
    n=8*pr.Nbase1
    print(spec,gfrac,freqs,M,[relerr(xb[i*n:(i+1)*n],xa[i*n:(i+1)*n]) for i in range(len(freqs))])
run(False,0.0,[150e6])
run(False,0.0,[145e6,150e6,155e6])
run(True,0.0,[150e6])
run(True,0.0,[145e6,150e6,155e6])
run(False,0.3,[145e6,150e6,155e6])
run(True,0.3,[145e6,150e6,155e6],M=4)
run(True,0.3,[145e6,150e6,155e6],M=3)
