"""Run under torchrun on >= 2 GPUs (tests/test_gpu_multi.py launches it): the cluster-sharded solve
against the single-GPU solve of the same problem — sagecal_b200.dist.verify_sharding: cost / gradient
to 1e-11, the LBFGS stage to 1e-6 on the Jones, two block-Jacobi SAGE sweeps (beta=1) against their
emulation by single-GPU solves of each block to 1e-7, the quality of the default solve, bit-identical
results on every rank."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sagecal_b200 import lib as blib, dist as sdist  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api = blib.load()
    stream = torch.cuda.Stream()
    api.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        rep = sdist.verify_sharding(api, rank, world)
    if rank == 0:
        print(json.dumps(rep))
        print("MGPU_CHECK", "OK" if rep["ok"] else "FAIL")
    api.lib.dirac_b200_nccl_finalize()
    dist.destroy_process_group()
    sys.exit(0 if rep["ok"] else 1)


if __name__ == "__main__":
    main()
