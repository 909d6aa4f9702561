"""Cluster-sharded calibration over the GPUs of one box.

One process per GPU.  Each rank owns a contiguous block of clusters — its slice of the
coherencies, by far the largest array — while the data, the residual and the Jones vector are
replicated.  All solver logic AND the collectives live in the C library: it calls ncclAllReduce on
its own stream through a communicator it owns (`dirac_b200_nccl_init`, csrc/comm.cu).  The host only
has to carry the 128-byte NCCL id from rank 0 to the others; here `torch.distributed` does that
(`init_nccl`).  A host-supplied callback (`make_allreduce`) is still accepted — the gloo CPU tests of
the plumbing use it.

Collectives per solve: ONE all-reduce per SAGE sweep ([residual delta | Jones delta | nerr] in one
message); in the LBFGS stage one all-reduce of the line model (three vectors, contiguous) per
iteration and one of the gradient.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .dirac_api import SkyModel, c_double_p, cptr, dptr

ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p)


def partition_clusters(M: int, world: int):
    """contiguous blocks of ceil(M/world) clusters (the reference's predict splits clusters over GPUs
    the same way, predict_withbeam_cuda.c:713-794); returns [(k0, k1)] per rank"""
    per = (M + world - 1) // world
    return [(min(r * per, M), min((r + 1) * per, M)) for r in range(world)]


class _CudaView:
    """zero-copy view of `count` doubles of device memory for torch.as_tensor"""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8",
                                         "data": (int(ptr), False), "version": 3, "strides": None}


def make_allreduce(device="cuda"):
    """returns (ctypes callback, keep-alive) summing a buffer of doubles over the default process
    group.  device='cpu' views host memory (gloo; used by the CPU tests of the plumbing)."""
    import torch
    import torch.distributed as dist

    views = {}  # (ptr, count) -> tensor view; (stream) -> ExternalStream: built once, the library
    streams = {}  # calls with the same few buffers thousands of times

    def _cb(ptr, count, stream, user):
        if device == "cuda":
            t = views.get((ptr, count))
            if t is None:
                t = views[(ptr, count)] = torch.as_tensor(_CudaView(ptr, count), device="cuda")
            st = streams.get(stream)
            if st is None:
                st = streams[stream] = torch.cuda.ExternalStream(int(stream))
            with torch.cuda.stream(st):
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
        else:
            a = np.ctypeslib.as_array(C.cast(ptr, c_double_p), shape=(int(count),))
            t = torch.from_numpy(a)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    cb = ALLREDUCE_FN(_cb)
    return cb


def init_nccl(api, rank: int, world: int):
    """create the library's own NCCL communicator; the id travels over the default process group"""
    import torch
    import torch.distributed as dist
    L = api.lib
    L.dirac_b200_nccl_unique_id.argtypes = [C.c_char_p]
    L.dirac_b200_nccl_init.argtypes = [C.c_int, C.c_int, C.c_char_p]
    buf = C.create_string_buffer(128)
    if rank == 0:
        if L.dirac_b200_nccl_unique_id(buf) != 0:
            raise RuntimeError("dirac_b200_nccl_unique_id failed")
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    ident = bytes(t.cpu().tolist())
    if L.dirac_b200_nccl_init(rank, world, ident) != 0:
        raise RuntimeError("dirac_b200_nccl_init failed")


class ShardedProblem:
    """this rank's shard of a solve interval, resident on its GPU"""

    def __init__(self, api, pr, barr, rank, world, beta=0.0, coh_local=None, use_callback=False):
        """pr: sagecal_b200.synth.Problem with ALL clusters (coh may be None when coh_local is
        given or generated on the device)"""
        self.api = api
        self.rank, self.world = rank, world
        self.k0, self.k1 = partition_clusters(pr.M, world)[rank]
        chunk_before = int(sum(pr.nchunk[: self.k0]))
        self.sky = SkyModel(pr.clusters[self.k0: self.k1], pr.N, p_base=8 * pr.N * chunk_before)
        self.npar = 8 * pr.N * pr.Mt
        self.n = 8 * pr.Nbase1
        if coh_local is None and pr.coh is not None:
            c = pr.coh.reshape(pr.Nbase1, pr.M, 4)[:, self.k0: self.k1, :]
            coh_local = np.ascontiguousarray(c).reshape(-1)
        L = api.lib
        L.dirac_b200_create_shard.restype = C.c_void_p
        L.dirac_b200_create_shard.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_int, C.c_int, C.c_longlong, c_double_p,
                                              c_double_p]
        L.dirac_b200_set_comm.argtypes = [C.c_void_p, C.c_int, C.c_int, ALLREDUCE_FN, C.c_void_p,
                                          C.c_int, C.c_int, C.c_double]
        self.h = L.dirac_b200_create_shard(pr.N, pr.Nbase, pr.tilesz, C.cast(barr, C.c_void_p),
                                           C.cast(self.sky.arr, C.c_void_p), self.sky.M,
                                           self.sky.Mt, self.npar,
                                           cptr(coh_local) if coh_local is not None else None,
                                           dptr(pr.x))
        L.dirac_b200_nccl_ready.restype = C.c_int
        if use_callback:
            self._cb = make_allreduce("cuda")
        else:
            if not L.dirac_b200_nccl_ready():
                init_nccl(api, rank, world)
            self._cb = None  # NULL: the library's own ncclAllReduce
        L.dirac_b200_set_comm(self.h, rank, world, self._cb, None, pr.M, self.k0, float(beta))

    def sagefit(self, pp, x_out=None, max_emiter=3, max_iter=2, max_lbfgs=10, lbfgs_m=7, linsolv=0,
                solver_mode=1, nulow=2.0, nuhigh=30.0, randomize=0):
        nu, r0, r1 = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0)
        rv = self.api.lib.dirac_b200_sagefit(self.h, dptr(pp),
                                             dptr(x_out) if x_out is not None else None,
                                             max_emiter, max_iter, max_lbfgs, lbfgs_m, linsolv,
                                             solver_mode, nulow, nuhigh, randomize, C.byref(nu),
                                             C.byref(r0), C.byref(r1))
        return rv, nu.value, r0.value, r1.value

    def close(self):
        if self.h:
            self.api.lib.dirac_b200_destroy(self.h)
            self.h = None
