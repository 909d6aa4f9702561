"""Consensus (ADMM) calibration over frequency subbands, one subband per GPU (BASELINE.json config 5).

Host-side driver of the C entry points `dirac_b200_consensus_*` / `dirac_b200_sagefit_admm`
(csrc/consensus.cu).  The reference spreads this over a master and slave processes
(src/MPI/sagecal_master.cpp:844-877, sagecal_slave.cpp:831-878); here every rank owns one subband and
the only exchange is ONE all-reduce of Npoly*8*N*Mt doubles per ADMM iteration, issued by the C
library on its stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .dirac_api import c_double_p, dptr


def basis(api, freqs, freq0, Npoly, ptype=1):
    """B[f, p] (setup_polynomials, consensus_poly.c:38)"""
    freqs = np.ascontiguousarray(freqs, dtype=np.float64)
    B = np.zeros((len(freqs), Npoly))
    L = api.lib
    L.dirac_b200_consensus_basis.argtypes = [c_double_p, C.c_int, C.c_int, c_double_p, C.c_double, C.c_int]
    if L.dirac_b200_consensus_basis(dptr(B.reshape(-1)), Npoly, len(freqs), dptr(freqs), float(freq0), ptype):
        raise ValueError("unknown polynomial type %r" % ptype)
    return B


def prod_inverse(api, B, rho):
    """Bi[k] = pinv(sum_f rho[f, k] B_f B_f^T) (find_prod_inverse_full, consensus_poly.c:465);
    rho: [Nf, M]"""
    Nf, Npoly = B.shape
    rho = np.ascontiguousarray(rho, dtype=np.float64)
    M = rho.shape[1]
    Bi = np.zeros((M, Npoly, Npoly))
    L = api.lib
    L.dirac_b200_consensus_prod_inverse.argtypes = [c_double_p, c_double_p, C.c_int, C.c_int, C.c_int, c_double_p]
    L.dirac_b200_consensus_prod_inverse(dptr(np.ascontiguousarray(B).reshape(-1)), dptr(Bi.reshape(-1)),
                                        Npoly, Nf, M, dptr(rho.reshape(-1)))
    return Bi


def step_numpy(J, Y, BZ, rho_i, Bf, Bi, clus_of, allreduce):
    """numpy restatement of one exchange (what dirac_b200_consensus_step does on the device), with
    `allreduce(z)` summing over the subbands in place.  Test infrastructure for the CPU (gloo) tests."""
    Y = Y + rho_i * J
    z = Bf[:, None] * Y[None, :]
    allreduce(z)
    cw = np.einsum("kpq,q->kp", Bi, Bf)            # c_k = Bi_k B_f
    bz = np.einsum("ip,pi->i", cw[clus_of], z)
    Ynew = Y - rho_i * bz
    return Ynew, bz, float(np.linalg.norm(J - bz)), float(np.linalg.norm(bz - BZ))


class ConsensusSubband:
    """this rank's subband of a consensus run on its GPU"""

    def __init__(self, api, dp, f_index, freqs, freq0, Npoly, rho, ptype=1):
        """dp: lib.DeviceProblem of this subband; rho: [M] regularisation per cluster (the same on
        every subband here); freqs: centre frequency of every subband"""
        self.api, self.dp = api, dp
        self.Npoly = Npoly
        self.B = basis(api, freqs, freq0, Npoly, ptype)
        self.rho = np.ascontiguousarray(rho, dtype=np.float64)
        self.Bi = prod_inverse(api, self.B, np.tile(self.rho, (len(freqs), 1)))
        self.Bf = np.ascontiguousarray(self.B[f_index])
        self.Y = np.zeros(dp.m)
        self.BZ = np.zeros(dp.m)
        L = api.lib
        L.dirac_b200_consensus_step.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                                c_double_p, c_double_p, C.c_int, c_double_p, c_double_p]
        L.dirac_b200_sagefit_admm.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                              c_double_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p,
                                              c_double_p]

    def exchange(self, J):
        pr, du = C.c_double(0.0), C.c_double(0.0)
        self.api.lib.dirac_b200_consensus_step(self.dp.h, dptr(J), dptr(self.Y), dptr(self.BZ),
                                               dptr(self.rho), dptr(self.Bf), dptr(self.Bi.reshape(-1)),
                                               self.Npoly, C.byref(pr), C.byref(du))
        return pr.value, du.value

    def jupdate(self, pp, max_emiter=1, max_iter=2, first=False, solver="lm"):
        """first ADMM iteration: plain calibration; later ones carry the consensus terms.
        solver: "rtr" = the reference's robust Riemannian trust-region J-update
        (dirac_b200_sagefit_admm_rtr), "lm" = this library's LM on the augmented cost"""
        if first:
            rv, _, r0, r1 = self.dp.sagefit(pp, None, max_emiter=max_emiter, max_iter=max_iter,
                                            max_lbfgs=0, solver_mode=1)
            return rv, r0, r1
        r0, r1 = C.c_double(0.0), C.c_double(0.0)
        if solver == "rtr":
            nu = C.c_double(0.0)
            L = self.api.lib
            L.dirac_b200_sagefit_admm_rtr.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p,
                                                      c_double_p, c_double_p, C.c_int, C.c_int,
                                                      C.c_double, C.c_double, C.c_int, c_double_p,
                                                      c_double_p, c_double_p]
            rv = L.dirac_b200_sagefit_admm_rtr(self.dp.h, dptr(pp), None, dptr(self.Y),
                                               dptr(self.BZ), dptr(self.rho), max_emiter, max_iter,
                                               2.0, 30.0, 0, C.byref(nu), C.byref(r0), C.byref(r1))
            return rv, r0.value, r1.value
        rv = self.api.lib.dirac_b200_sagefit_admm(self.dp.h, dptr(pp), None, dptr(self.Y), dptr(self.BZ),
                                                  dptr(self.rho), max_emiter, max_iter, 0, 0,
                                                  C.byref(r0), C.byref(r1))
        return rv, r0.value, r1.value

    def run(self, pp, admm_iters=5, max_emiter=1, max_iter=2, solver="lm"):
        """ADMM loop (sagecal_slave.cpp:700-900 without the master); returns per-iteration
        (res_0, res_1, primal, dual)"""
        hist = []
        for it in range(admm_iters):
            rv, r0, r1 = self.jupdate(pp, max_emiter, max_iter, first=(it == 0), solver=solver)
            pr, du = self.exchange(pp)
            hist.append((r0, r1, pr, du))
        return hist
