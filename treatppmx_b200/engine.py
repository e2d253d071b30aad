"""Host-side handles over the C ABI (include/tppmx.h): a device context and a batch of chains.
Thin: argument marshalling and error mapping only; every number is computed by the CUDA
library."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._abi import (MAXD, Dataset, Dims, HostDataset, HostShared, HostState, Model, c_dp, c_ip)


class DeviceArray:
    """A library-owned device buffer exposed through __cuda_array_interface__ (zero-copy view for
    torch.as_tensor(..., device="cuda") so NCCL can gather it without a host round trip)."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.ptr, self.shape, self.typestr, self._owner = ptr, tuple(int(x) for x in shape), typestr, owner

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.typestr, "data": (self.ptr, False), "version": 2}


class Context:
    def __init__(self, device: int = 0):
        self.L = _lib.load()
        self.h = C.c_void_p()
        rc = self.L.tppmx_create(C.byref(self.h), device)
        if rc != 0:
            raise _lib.TppmxError(rc, "tppmx_create failed (no usable CUDA device?)")
        self.device = device

    def err(self) -> str:
        return self.L.tppmx_last_error(self.h).decode()

    def check(self, rc: int):
        if rc != 0:
            raise _lib.TppmxError(rc, self.err())

    def set_interrupt_flag(self, flag, check_every: int = 64):
        """flag: a one-element int32 numpy array the caller (another thread, a signal handler) sets to
        non-zero to stop tppmx_batch_run / tppmx_dm_ppmx_ct (TppmxError code 5); None switches it off."""
        self._flag = flag   # keep the buffer alive
        self.check(self.L.tppmx_set_interrupt_flag(self.h, flag.ctypes.data_as(c_ip) if flag is not None else c_ip(),
                                                   check_every))

    def probe_math(self, which: str, x):
        """Device fp64 log / exp / lgamma of the fast sweep kernel on an array (test hook)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        self.check(self.L.tppmx_probe_math(self.h, {"log": 0, "exp": 1, "lgamma": 2, "logtab": 3, "lgammatab": 4}[which],
                                           x.ctypes.data_as(c_dp), x.size, out.ctypes.data_as(c_dp)))
        return out

    def close(self):
        if self.h:
            self.L.tppmx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """n_chains independent chains over n_datasets replicate datasets on one GPU."""

    def __init__(self, ctx: Context, model: Model, dims: Dims, shared: HostShared,
                 datasets: list[HostDataset], chain_dataset, chain_id=None):
        self.ctx, self.model, self.dims, self.shared, self.datasets = ctx, model, dims, shared, datasets
        self.chain_dataset = np.ascontiguousarray(chain_dataset, dtype=np.int32)
        self.n_chains = int(self.chain_dataset.size)
        self.chain_id = (np.arange(self.n_chains, dtype=np.int32) if chain_id is None
                         else np.ascontiguousarray(chain_id, dtype=np.int32))
        arr = (Dataset * len(datasets))(*[d.cstruct() for d in datasets])
        sh = shared.cstruct()
        self.h = C.c_void_p()
        ctx.check(ctx.L.tppmx_batch_create(
            ctx.h, C.byref(model), C.byref(dims), C.byref(sh), len(datasets), arr, self.n_chains,
            self.chain_dataset.ctypes.data_as(c_ip), self.chain_id.ctypes.data_as(c_ip),
            C.byref(self.h)))

    def close(self):
        if self.h:
            self.ctx.L.tppmx_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self, seed: int, init_labels, init_nclu, initbeta):
        """init_labels: [n_datasets] x (T x ntmax F-order) or one array broadcast to all."""
        nd, T, nt = len(self.datasets), self.dims.T, self.dims.ntmax
        il = np.asarray(init_labels, dtype=np.int32)
        if il.ndim == 2:
            il = np.broadcast_to(il, (nd, T, nt))
        il = np.ascontiguousarray(il.transpose(0, 2, 1)).reshape(nd, T * nt)   # each dataset's T x ntmax, column-major
        ic = np.asarray(init_nclu, dtype=np.int32)
        if ic.ndim == 1:
            ic = np.broadcast_to(ic, (nd, T))
        ic = np.ascontiguousarray(ic)
        ib = np.ascontiguousarray(initbeta, dtype=np.float64)
        self.ctx.check(self.ctx.L.tppmx_batch_init(
            self.h, seed, il.ctypes.data_as(c_ip), ic.ctypes.data_as(c_ip), ib.ctypes.data_as(c_dp)))

    def set_state(self, chain: int, st: HostState):
        cs = st.cstruct()
        self.ctx.check(self.ctx.L.tppmx_batch_set_state(self.h, chain, C.byref(cs)))

    def get_state(self, chain: int) -> HostState:
        st = HostState(self.dims, self.model.CC)
        cs = st.cstruct()
        self.ctx.check(self.ctx.L.tppmx_batch_get_state(self.h, chain, C.byref(cs)))
        return st

    def get_partitions(self):
        """labels [n_chains, T, ntmax] (1-based) and nclu [n_chains, T] of every chain."""
        d = self.dims
        lab = np.zeros((self.n_chains, d.ntmax, d.T), dtype=np.int32)  # C-order == col-major T x ntmax
        ncl = np.zeros((self.n_chains, d.T), dtype=np.int32)
        self.ctx.check(self.ctx.L.tppmx_batch_get_partitions(
            self.h, lab.ctypes.data_as(c_ip), ncl.ctypes.data_as(c_ip)))
        return lab.transpose(0, 2, 1), ncl

    def score_patient(self, chain: int, arm: int, it: int, seed: int = 0, iter_: int = 0):
        n = self.dims.ntmax + self.model.CC
        logw, prob, K = np.zeros(n), np.zeros(n), C.c_int32(0)
        self.ctx.check(self.ctx.L.tppmx_score_patient(
            self.h, chain, arm, it, seed, iter_, logw.ctypes.data_as(c_dp),
            prob.ctypes.data_as(c_dp), C.byref(K)))
        k = K.value + self.model.CC
        return logw[:k], prob[:k], K.value

    def sweep(self, seed: int, iter0: int, n_sweeps: int):
        self.ctx.check(self.ctx.L.tppmx_batch_sweep(self.h, seed, iter0, n_sweeps))

    def sweep_probe(self, seed: int, iter_: int, chain: int):
        """One sweep of every chain through the probe instantiation of the fast sweep kernel: returns
        (prob [T, ntmax, W], K [T, ntmax]) of `chain` -- each step's normalised allocation
        probabilities and occupied-cluster count (-1 where the fast kernel did not run the step)."""
        d = self.dims
        W = min(d.kcap if d.kcap > 0 else d.ntmax, 256) + self.model.CC   # >= the staged capacity of either kernel
        prob = np.zeros((d.T, d.ntmax, W))
        K = np.zeros((d.T, d.ntmax), dtype=np.int32)
        self.ctx.check(self.ctx.L.tppmx_batch_sweep_probe(self.h, seed, iter_, chain, W, prob.ctypes.data_as(c_dp),
                                                          K.ctypes.data_as(c_ip)))
        return prob, K

    # kernel selection (include/tppmx.h: TPPMX_OPT_*, TPPMX_INFO_*)
    IMPL = {"auto": 0, "generic": 1, "fast": 2}

    def set_sweep_impl(self, impl: str = "auto", lanes_per_unit: int = 0, staged_clusters: int = 0, big_threads: int = 0):
        self.ctx.check(self.ctx.L.tppmx_batch_set_option(self.h, 1, self.IMPL[impl]))
        self.ctx.check(self.ctx.L.tppmx_batch_set_option(self.h, 2, lanes_per_unit))
        if staged_clusters:
            self.ctx.check(self.ctx.L.tppmx_batch_set_option(self.h, 3, staged_clusters))
        if big_threads:
            self.ctx.check(self.ctx.L.tppmx_batch_set_option(self.h, 4, big_threads))

    def set_chain_groups(self, groups: int = 0):
        """run(): split the chains into `groups` contiguous groups iterating on their own streams (0 = automatic);
        results do not depend on it (TPPMX_OPT_CHAIN_GROUPS)"""
        self.ctx.check(self.ctx.L.tppmx_batch_set_option(self.h, 5, groups))

    def last_chain_groups(self) -> int:
        a = C.c_int32(0)
        self.ctx.check(self.ctx.L.tppmx_batch_get_info(self.h, 3, C.byref(a)))
        return a.value

    def last_sweep_info(self):
        """(kernel used by the last sweep: 'generic'|'fast', units handed over to generic)"""
        a, c = C.c_int32(0), C.c_int32(0)
        self.ctx.check(self.ctx.L.tppmx_batch_get_info(self.h, 1, C.byref(a)))
        self.ctx.check(self.ctx.L.tppmx_batch_get_info(self.h, 2, C.byref(c)))
        return {1: "generic", 2: "fast"}.get(a.value, "none"), c.value

    def last_timing(self):
        ms, n = C.c_double(0), C.c_int32(0)
        self.ctx.check(self.ctx.L.tppmx_batch_last_timing(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- full iterations on the device (sweep + the remaining updates) ---------------------------
    def update(self, seed: int, iter_: int):
        self.ctx.check(self.ctx.L.tppmx_batch_update(self.h, seed, iter_))

    def run(self, seed: int, iter0: int, n_iter: int, burn: int = 0, thin: int = 1, psm: bool = False,
            predictive: bool = False):
        """n_iter full iterations; saved iterations ((l > burn-1) and (l+1) % thin == 0) feed the
        device summaries."""
        self.ctx.check(self.ctx.L.tppmx_batch_run(self.h, seed, iter0, n_iter, burn, thin, int(psm), int(predictive)))

    def get_acceptance(self, chain: int):
        d = self.dims
        ea = np.zeros((d.T, d.ntmax), dtype=np.int32, order="F")
        ba = np.zeros((max(d.Q, 1), d.D), dtype=np.int32, order="F")
        self.ctx.check(self.ctx.L.tppmx_batch_get_acceptance(self.h, chain, ea.ctypes.data_as(c_ip), ba.ctypes.data_as(c_ip)))
        return ea, ba

    # ---- posterior summaries accumulated on the device (include/tppmx.h) -------------------
    def reset_summaries(self):
        self.ctx.check(self.ctx.L.tppmx_batch_reset_summaries(self.h))

    def accumulate(self, seed: int, iter_: int, psm: bool = True, predictive: bool = True):
        self.ctx.check(self.ctx.L.tppmx_batch_accumulate(self.h, seed, iter_, int(psm), int(predictive)))

    def summaries(self, util_w=None, want=("psm", "pi_mean", "utility", "best_arm")):
        """Host copies: psm [nd, T, ntmax, ntmax], pi_mean [nd, npred, D, T], utility [nd, T, npred],
        best_arm [nd, npred]."""
        d, nd = self.dims, len(self.datasets)
        uw = np.zeros(MAXD) if util_w is None else np.concatenate([np.asarray(util_w, float), np.zeros(MAXD)])[:MAXD]
        uw = np.ascontiguousarray(uw)
        psm = np.zeros((nd, d.T, d.ntmax, d.ntmax)) if "psm" in want else None
        pim = np.zeros((nd, d.T, d.D, d.npred)) if "pi_mean" in want and d.npred > 0 else None
        ut = np.zeros((nd, d.T, d.npred)) if "utility" in want and d.npred > 0 else None
        ba = np.zeros((nd, d.npred), dtype=np.int32) if "best_arm" in want and d.npred > 0 else None
        f = lambda a, t: a.ctypes.data_as(t) if a is not None else t()
        self.ctx.check(self.ctx.L.tppmx_batch_summaries(self.h, uw.ctypes.data_as(c_dp), f(psm, c_dp), f(pim, c_dp),
                                                        f(ut, c_dp), f(ba, c_ip)))
        return dict(psm=psm, pi_mean=None if pim is None else pim.transpose(0, 3, 2, 1), utility=ut, best_arm=ba)

    def npc(self, trtsgn, outcome):
        """npc of R/countUT.R:26-53 per replicate dataset, from the predictive means of the last
        summaries() call: trtsgn / outcome are [n_datasets, npred], 0-based."""
        nd, npred = len(self.datasets), self.dims.npred
        t = np.ascontiguousarray(np.asarray(trtsgn, dtype=np.int32).reshape(nd, npred))
        o = np.ascontiguousarray(np.asarray(outcome, dtype=np.int32).reshape(nd, npred))
        out = np.zeros(nd, dtype=np.int32)
        self.ctx.check(self.ctx.L.tppmx_batch_npc(self.h, t.ctypes.data_as(c_ip), o.ctypes.data_as(c_ip),
                                                  out.ctypes.data_as(c_ip)))
        return out

    def point_estimate(self, loss: str = "VI", linkage: str = "average", max_k: int = 0):
        """Point-estimate partition per (dataset, arm) from the co-clustering matrix of the last
        summaries() call (minVI / minbinder.ext, method "avg" | "comp"): labels [nd, T, ntmax] 1-based
        (0 beyond the arm), k [nd, T], loss [nd, T]."""
        nd, d = len(self.datasets), self.dims
        lab = np.zeros((nd, d.ntmax, d.T), dtype=np.int32)      # C-order == col-major T x ntmax per dataset
        k = np.zeros((nd, d.T), dtype=np.int32)
        ls = np.zeros((nd, d.T))
        self.ctx.check(self.ctx.L.tppmx_batch_point_estimate(
            self.h, {"binder": 0, "VI": 1}[loss], {"average": 0, "complete": 1}[linkage], max_k,
            lab.ctypes.data_as(c_ip), k.ctypes.data_as(c_ip), ls.ctypes.data_as(c_dp)))
        return lab.transpose(0, 2, 1), k, ls

    def summaries_device(self, util_w=None):
        """The same arrays left on the device as DeviceArray views (C-order shapes as in summaries(),
        pi_mean as [nd, T, D, npred])."""
        d, nd = self.dims, len(self.datasets)
        uw = np.zeros(MAXD) if util_w is None else np.concatenate([np.asarray(util_w, float), np.zeros(MAXD)])[:MAXD]
        uw = np.ascontiguousarray(uw)
        p1, p2, p3, p4 = c_dp(), c_dp(), c_dp(), c_ip()
        self.ctx.check(self.ctx.L.tppmx_batch_summaries_device(self.h, uw.ctypes.data_as(c_dp), C.byref(p1), C.byref(p2),
                                                               C.byref(p3), C.byref(p4)))
        addr = lambda p: C.cast(p, C.c_void_p).value
        out = {}
        if addr(p1):
            out["psm"] = DeviceArray(addr(p1), (nd, d.T, d.ntmax, d.ntmax), "<f8", self)
        if d.npred > 0:
            out["pi_mean"] = DeviceArray(addr(p2), (nd, d.T, d.D, d.npred), "<f8", self)
            out["utility"] = DeviceArray(addr(p3), (nd, d.T, d.npred), "<f8", self)
            out["best_arm"] = DeviceArray(addr(p4), (nd, d.npred), "<i4", self)
        return out

    def predict(self, seed: int, iter_: int, want=("pipred", "ypred", "predclass")):
        d, nc = self.dims, self.n_chains
        pipred = np.zeros((nc, d.T, d.D, d.npred)) if "pipred" in want else None
        ypred = np.zeros((nc, d.T, d.D, d.npred)) if "ypred" in want else None
        pc = np.zeros((nc, d.T, d.npred), dtype=np.int32) if "predclass" in want else None
        f = lambda a, t: a.ctypes.data_as(t) if a is not None else t()
        self.ctx.check(self.ctx.L.tppmx_batch_predict(
            self.h, seed, iter_, f(pipred, c_dp), f(ypred, c_dp), f(pc, c_ip)))
        # C-order [chain][T][D][npred] is the col-major npred x D x T cube of the reference
        tr = lambda a: None if a is None else a.transpose(0, 3, 2, 1)
        return tr(pipred), tr(ypred), None if pc is None else pc.transpose(0, 2, 1)
