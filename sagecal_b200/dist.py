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
            self._cb = C.cast(None, ALLREDUCE_FN)  # NULL: the library's own ncclAllReduce
        L.dirac_b200_set_comm(self.h, rank, world, self._cb, None, pr.M, self.k0, float(beta))

    def precalculate(self, u, v, w, freq0, fdelta, uvmin=0.0, uvmax=1e9):
        """coherencies of this rank's clusters generated on the device (dirac_b200_precalculate)"""
        L = self.api.lib
        L.dirac_b200_precalculate.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p,
                                              C.c_void_p, C.c_double, C.c_double, C.c_double,
                                              C.c_double, C.c_void_p]
        L.dirac_b200_precalculate(self.h, dptr(u), dptr(v), dptr(w), C.cast(self.sky.arr, C.c_void_p),
                                  freq0, fdelta, uvmin, uvmax, None)

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


def _models(pr, pp):
    """per-cluster model visibilities J_p C J_q^H (the gauge-invariant content of a solution)"""
    from . import synth
    out = []
    nrow = pr.Nbase1
    c = pr.coh.reshape(nrow, pr.M, 4)
    off = 0
    for k in range(pr.M):
        nch = pr.nchunk[k]
        blk = pp[off: off + nch * 8 * pr.N]
        out.append(synth.apply_jones(np.ascontiguousarray(c[:, k: k + 1, :]).reshape(-1), blk,
                                     pr.sta1, pr.sta2, pr.N, [nch], pr.flag))
        off += nch * 8 * pr.N
    return out


def verify_sharding(api, rank, world, N=16, M=None, tilesz=10, seed=77):
    """Correctness of the cluster-sharded path against the single-GPU path on a small problem (every
    rank computes the single-GPU yardsticks itself).  Returns a dict of max relative errors; `ok`
    applies the tolerances written next to each entry.
      cost, grad     same function, only the order of the sums differs                    1e-11
      lbfgs_only     LBFGS stage alone (max_emiter=0): the same iteration on the same cost, every
                     decision of the line search is shared, so the Jones agree              1e-6
      jacobi         two SAGE sweeps with hidden-data weight beta=1 are block Jacobi across the
                     ranks: rank g solves its clusters (Gauss-Seidel among them) against
                     x - sum_{k not in g} model_k(p_start).  Exactly that is emulated with
                     single-GPU solves of each block on substituted data; Jones and final
                     residual of the sharded run must reproduce it                          1e-7
      quality        full solve with the default weight beta=1/world: the residual the sharded
                     solution leaves is within 25 % of the sequential one after 4 sweeps (the iterates of block
                     Jacobi and Gauss-Seidel differ; per-cluster models of weak clusters are not
                     identifiable at this size, so only the residual is compared)
      identical      every rank holds bit-identical Jones and residual"""
    import torch
    import torch.distributed as dist
    from . import lib as blib, synth
    from .dirac_api import make_barr
    M = M or 4 * world
    pr = synth.make_problem(N=N, M=M, tilesz=tilesz, seed=seed, kmean=1.0)
    barr = make_barr(pr.sta1, pr.sta2, pr.flag)
    sky = SkyModel(pr.clusters, pr.N)
    rng = np.random.default_rng(seed + 1)
    pprobe = pr.pp0 + 0.1 * rng.normal(0, 1, pr.pp0.shape)
    rep = {}
    L = api.lib
    n8 = 8 * pr.N

    def rel(a, b):
        return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / (np.max(np.abs(b)) + 1e-300))

    dp = blib.DeviceProblem(api, pr.N, pr.Nbase, pr.tilesz, barr, sky, pr.coh, pr.x)
    sp = ShardedProblem(api, pr, barr, rank, world)
    # cost and gradient through the thin layer (sharded handle: partial models / gradients summed)
    c1, g1 = dp.cost(pprobe), dp.grad(pprobe)
    L.dirac_b200_predict.restype = C.c_double
    cs = L.dirac_b200_predict(sp.h, dptr(pprobe), None, 0, 1, 0.0)
    gs = np.zeros_like(pprobe)
    L.dirac_b200_grad(sp.h, dptr(pprobe), dptr(gs), 0, 0.0)
    rep["cost"] = abs(cs - c1) / c1
    rep["grad"] = rel(gs, g1)
    # LBFGS stage alone
    kw = dict(max_emiter=0, max_iter=0, max_lbfgs=12, lbfgs_m=7, solver_mode=1)
    p1, ps = pprobe.copy(), pprobe.copy()
    r1 = dp.sagefit(p1, None, **kw)
    rs = sp.sagefit(ps, None, **kw)
    rep["lbfgs_only_jones"] = rel(ps, p1)
    rep["lbfgs_only_res1"] = abs(rs[3] - r1[3]) / r1[3]
    # full solve, default beta: quality of the solution + identical results on every rank
    kw = dict(max_emiter=4, max_iter=3, max_lbfgs=10, lbfgs_m=7, solver_mode=1)
    p1, ps = pr.pp0.copy(), pr.pp0.copy()
    xs = np.zeros_like(pr.x)
    r1 = dp.sagefit(p1, None, **kw)
    rs = sp.sagefit(ps, xs, **kw)
    rep["res0"] = abs(rs[2] - r1[2]) / r1[2]
    rep["quality_res1_sharded_over_single"] = rs[3] / r1[3]
    rep["rv"] = [r1[0], rs[0]]
    sp.close()
    t = torch.from_numpy(np.concatenate([ps, xs, [rs[2], rs[3]]])).cuda()
    t0 = t.clone()
    dist.broadcast(t0, 0)
    same = torch.tensor([1.0 if torch.equal(t, t0) else 0.0], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    rep["identical_on_ranks"] = bool(same.item() > 0.5)
    # block-Jacobi sweeps (beta = 1) against their single-GPU emulation
    sweeps, kwj = 2, dict(max_iter=3, max_lbfgs=0, lbfgs_m=7, solver_mode=1)
    spj = ShardedProblem(api, pr, barr, rank, world, beta=1.0)
    pj = pr.pp0.copy()
    rj = spj.sagefit(pj, None, max_emiter=sweeps, **kwj)
    spj.close()
    parts = partition_clusters(pr.M, world)
    cohk = pr.coh.reshape(pr.Nbase1, pr.M, 4)
    pe = pr.pp0.copy()
    for _ in range(sweeps):
        models = _models(pr, pe)
        total = np.sum(models, axis=0)
        nxt = pe.copy()
        for (k0, k1) in parts:
            if k1 <= k0:
                continue
            # this block's data: everything the other blocks explain at the sweep's start removed
            xb = pr.x - (total - np.sum(models[k0:k1], axis=0))
            skyb = SkyModel(pr.clusters[k0:k1], pr.N)
            cohb = np.ascontiguousarray(cohk[:, k0:k1, :]).reshape(-1)
            with blib.DeviceProblem(api, pr.N, pr.Nbase, pr.tilesz, barr, skyb, cohb, xb) as db:
                o0 = int(sum(pr.nchunk[:k0])) * n8
                o1 = int(sum(pr.nchunk[:k1])) * n8
                pb = pe[o0:o1].copy()
                db.sagefit(pb, None, max_emiter=1, **kwj)
                nxt[o0:o1] = pb
        pe = nxt
    rep["jacobi_jones"] = rel(pj, pe)
    res_e = np.sqrt(dp.cost(pe)) / (8.0 * pr.Nbase1)
    rep["jacobi_res1"] = abs(rj[3] - res_e) / res_e
    dp.close()
    rep["world"] = world
    rep["ok"] = bool(rep["cost"] < 1e-11 and rep["grad"] < 1e-11 and rep["lbfgs_only_jones"] < 1e-6
                     and rep["res0"] < 1e-10 and rep["jacobi_jones"] < 1e-7 and rep["jacobi_res1"] < 1e-7
                     and rep["quality_res1_sharded_over_single"] < 1.25 and rs[0] == 0
                     and rep["identical_on_ranks"])
    return rep
