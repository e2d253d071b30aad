"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs.
Never imported by treatppmx_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from treatppmx_b200._abi import (Dataset, Dims, HostDataset, HostShared, HostState, Model,
                                 OracleProblem, Shared, State, c_dp, c_ip)

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(_ROOT, "oracle", f) for f in os.listdir(os.path.join(_ROOT, "oracle"))
            if f.endswith((".c", ".h"))] + [os.path.join(_ROOT, "include", "tppmx.h")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "-B"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        d, i, u32, u64 = C.c_double, C.c_int, C.c_uint32, C.c_uint64
        L.oracle_dnorm_log.restype = d
        L.oracle_dnorm_log.argtypes = [d, d, d]
        L.oracle_dinvgamma.restype = d
        L.oracle_dinvgamma.argtypes = [d, d, d, i]
        L.oracle_dN_IG.restype = d
        L.oracle_dN_IG.argtypes = [d, d, d, d, d, d, i]
        L.oracle_gsimconNN.restype = d
        L.oracle_gsimconNN.argtypes = [d, d, d, d, d, i, i, i]
        L.oracle_gsimconNNIG.restype = d
        L.oracle_gsimconNNIG.argtypes = [d, d, d, d, d, d, i, i, i]
        L.oracle_gsimcatDM.restype = d
        L.oracle_gsimcatDM.argtypes = [c_dp, c_dp, i, i, i]
        L.oracle_rng_u2.argtypes = [u64, u32, u32, u32, u32, u32, u32, c_dp]
        L.oracle_rng_n2.argtypes = [u64, u32, u32, u32, u32, u32, u32, c_dp]
        L.oracle_rng_gamma.restype = d
        L.oracle_rng_gamma.argtypes = [u64, u32, u32, u32, u32, u32, d]
        P = C.POINTER(OracleProblem)
        S = C.POINTER(State)
        L.oracle_init_state.argtypes = [P, S, u64, u32, c_ip, c_ip, c_dp]
        L.oracle_score_patient.argtypes = [P, S, C.c_int32, C.c_int32, u64, u32, C.c_int32,
                                           c_dp, c_dp, c_ip]
        L.oracle_sweep.argtypes = [P, S, u64, u32, C.c_int32, C.c_int32]
        L.oracle_predict.argtypes = [P, S, u64, u32, C.c_int32, c_dp, c_dp, c_ip]
        L.oracle_sweep_probe.argtypes = [P, S, u64, u32, C.c_int32, C.c_int32, c_dp, c_ip]
        i32p = c_ip
        L.oracle_update_iter.argtypes = [P, S, u64, u32, C.c_int32, i32p, i32p]
        L.oracle_iterate.argtypes = [P, S, u64, u32, C.c_int32, C.c_int32, i32p, i32p]
        L.oracle_insample.argtypes = [P, S, u64, u32, C.c_int32, c_dp, c_dp, c_dp]
        L.oracle_gammainc.restype = d
        L.oracle_gammainc.argtypes = [d, d]
        L.oracle_gammaincinv.restype = d
        L.oracle_gammaincinv.argtypes = [d, d]
        L.oracle_dmvnrm_log.restype = d
        L.oracle_dmvnrm_log.argtypes = [c_dp, c_dp, c_dp, i]
        L.oracle_wishart.argtypes = [u64, u32, C.c_int32, c_dp, d, i, c_dp]
        _lib = L
    return _lib


class Problem:
    """Read-only inputs of one chain, kept alive together with their numpy buffers."""

    def __init__(self, model: Model, dims: Dims, shared: HostShared, data: HostDataset):
        self.model, self.dims, self.shared, self.data = model, dims, shared, data
        self.c = OracleProblem()
        self.c.model = model
        self.c.dims = dims
        self.c.shared = shared.cstruct()
        self.c.data = data.cstruct()

    def new_state(self) -> HostState:
        return HostState(self.dims, self.model.CC)

    def init_state(self, seed, chain_id, init_labels, init_nclu, initbeta) -> HostState:
        s = self.new_state()
        cs = s.cstruct()
        il = np.asfortranarray(init_labels, dtype=np.int32)
        ic = np.ascontiguousarray(init_nclu, dtype=np.int32)
        ib = np.ascontiguousarray(initbeta, dtype=np.float64)
        rc = lib().oracle_init_state(C.byref(self.c), C.byref(cs), seed, chain_id,
                                     il.ctypes.data_as(c_ip), ic.ctypes.data_as(c_ip),
                                     ib.ctypes.data_as(c_dp))
        assert rc == 0, rc
        return s

    def score_patient(self, s: HostState, arm, it, seed=0, chain_id=0, iter_=0):
        n = self.dims.ntmax + self.model.CC
        logw = np.zeros(n)
        prob = np.zeros(n)
        K = C.c_int32(0)
        cs = s.cstruct()
        rc = lib().oracle_score_patient(C.byref(self.c), C.byref(cs), arm, it, seed, chain_id,
                                        iter_, logw.ctypes.data_as(c_dp),
                                        prob.ctypes.data_as(c_dp), C.byref(K))
        assert rc == 0, rc
        k = K.value + self.model.CC
        return logw[:k], prob[:k], K.value

    def sweep(self, s: HostState, seed, chain_id, iter0, n_sweeps):
        cs = s.cstruct()
        rc = lib().oracle_sweep(C.byref(self.c), C.byref(cs), seed, chain_id, iter0, n_sweeps)
        assert rc == 0, rc

    def sweep_probe(self, s: HostState, seed, chain_id, iter_, stride):
        """One sweep recording every step's allocation probabilities: (prob [T, ntmax, stride], K [T, ntmax])."""
        d = self.dims
        prob = np.zeros((d.T, d.ntmax, stride))
        K = np.full((d.T, d.ntmax), -1, dtype=np.int32)
        cs = s.cstruct()
        rc = lib().oracle_sweep_probe(C.byref(self.c), C.byref(cs), seed, chain_id, iter_, stride,
                                      prob.ctypes.data_as(c_dp), K.ctypes.data_as(c_ip))
        assert rc == 0, rc
        return prob, K

    def new_acc(self):
        d = self.dims
        return (np.zeros((d.T, d.ntmax), dtype=np.int32, order="F"),
                np.zeros((max(d.Q, 1), d.D), dtype=np.int32, order="F"))

    def update_iter(self, s: HostState, seed, chain_id, iter_, acc=None):
        """Everything of iteration `iter_` after the sweep (src/dm_ppmx_ct.cpp:866-1055)."""
        cs = s.cstruct()
        ea, ba = (acc if acc is not None else (None, None))
        f = lambda a: a.ctypes.data_as(c_ip) if a is not None else c_ip()
        rc = lib().oracle_update_iter(C.byref(self.c), C.byref(cs), seed, chain_id, iter_, f(ea), f(ba))
        assert rc == 0, rc

    def iterate(self, s: HostState, seed, chain_id, iter0, n_iter, acc=None):
        """n_iter full iterations: sweep + updates."""
        cs = s.cstruct()
        ea, ba = (acc if acc is not None else (None, None))
        f = lambda a: a.ctypes.data_as(c_ip) if a is not None else c_ip()
        rc = lib().oracle_iterate(C.byref(self.c), C.byref(cs), seed, chain_id, iter0, n_iter, f(ea), f(ba))
        assert rc == 0, rc

    def insample(self, s: HostState, seed, chain_id, iter_):
        """In-sample fit of a saved iteration (:1060-1075): pi, ispred (nobs x D), like (nobs)."""
        d = self.dims
        pi = np.zeros((d.nobs, d.D), order="F")
        isp = np.zeros((d.nobs, d.D), order="F")
        like = np.zeros(d.nobs)
        cs = s.cstruct()
        rc = lib().oracle_insample(C.byref(self.c), C.byref(cs), seed, chain_id, iter_, pi.ctypes.data_as(c_dp),
                                   isp.ctypes.data_as(c_dp), like.ctypes.data_as(c_dp))
        assert rc == 0, rc
        return pi, isp, like

    def predict(self, s: HostState, seed, chain_id, iter_):
        d = self.dims
        pipred = np.zeros((d.npred, d.D, d.T), order="F")
        ypred = np.zeros((d.npred, d.D, d.T), order="F")
        pc = np.zeros((d.npred, d.T), dtype=np.int32, order="F")
        cs = s.cstruct()
        rc = lib().oracle_predict(C.byref(self.c), C.byref(cs), seed, chain_id, iter_,
                                  pipred.ctypes.data_as(c_dp), ypred.ctypes.data_as(c_dp),
                                  pc.ctypes.data_as(c_ip))
        assert rc == 0, rc
        return pipred, ypred, pc
