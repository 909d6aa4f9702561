"""CPU (gloo, world_size 2): the host-side plumbing of the cluster-sharded solve — partitioning and
the all-reduce callback the C library is given (here on host memory; NCCL on the GPU box)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sagecal_b200 import dist as sdist


def test_partition_clusters():
    assert sdist.partition_clusters(64, 8) == [(8 * r, 8 * r + 8) for r in range(8)]
    assert sdist.partition_clusters(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert sdist.partition_clusters(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for M, W in ((256, 8), (7, 3), (1, 2)):
        parts = sdist.partition_clusters(M, W)
        assert parts[0][0] == 0 and parts[-1][1] == M
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cb = sdist.make_allreduce("cpu")
    # what the C library does: hand over a raw pointer, a count and a stream handle
    buf = np.arange(1000, dtype=np.float64) * (rank + 1)
    cb(buf.ctypes.data_as(C.c_void_p), 1000, None, None)
    want = np.arange(1000, dtype=np.float64) * sum(range(1, world + 1))
    ok = np.array_equal(buf, want)
    # the sharded bookkeeping: every rank owns a block, zero elsewhere, the sum is the full vector
    k0, k1 = sdist.partition_clusters(10, world)[rank]
    v = np.zeros(10)
    v[k0:k1] = np.arange(k0, k1) + 1.0
    cb(v.ctypes.data_as(C.c_void_p), 10, None, None)
    ok = ok and np.array_equal(v, np.arange(10) + 1.0)
    out[rank] = 1 if ok else 0
    dist.destroy_process_group()


def test_allreduce_callback_gloo_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: 1, 1: 1}
