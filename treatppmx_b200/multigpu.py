"""Multi-GPU plumbing for the path: shard replicate datasets (and their chains) across ranks,
run independently, gather the per-dataset posterior summaries at the end.

The path shards by independent units: chains never communicate, so there is no
data-path collective — only this final gather.  One process per GPU; `torch.distributed` is the
plumbing (backend "nccl" over NVLink on a GPU box, "gloo" in the CPU tests)."""
from __future__ import annotations

import numpy as np


def pin_to_gpu_numa_node(device: int) -> dict:
    """Restricts this process (and the threads it starts later) to the CPUs of the NUMA node the GPU hangs
    off, so that the pinned staging buffers are first-touched next to the device and the host threads that
    fill them run there (one process per GPU: eight unpinned processes share the box's memory channels
    across sockets).  Reads /sys only; returns what it did ({"node": n, "cpus": k} or {"skipped": why})."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device)
        pci = f"{getattr(bus, 'pci_domain_id', 0):04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read())
        if node < 0:
            return {"skipped": "no NUMA node reported for " + pci}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        mine = cpus & os.sched_getaffinity(0)
        if len(mine) < 2:
            return {"skipped": f"node {node} has {len(mine)} of this process's CPUs"}
        os.sched_setaffinity(0, mine)
        return {"node": node, "cpus": len(mine)}
    except Exception as e:   # never fatal: the run is merely not pinned
        return {"skipped": repr(e)}


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` datasets owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_chains(chain_dataset, world: int, rank: int, n_datasets: int):
    """Datasets [lo, hi) go to `rank`; returns (lo, hi, indices of the chains that follow them,
    their dataset index re-based to the shard)."""
    lo, hi = shard_range(n_datasets, world, rank)
    cd = np.asarray(chain_dataset)
    idx = np.nonzero((cd >= lo) & (cd < hi))[0]
    return lo, hi, idx, (cd[idx] - lo).astype(np.int32)


def gather_summaries(local: dict, n_datasets_total: int, group=None, dst=None):
    """All-gathers per-dataset summaries sharded with shard_range.

    local: name -> torch tensor whose leading dimension is the rank's number of datasets (CUDA
    tensors for nccl, CPU tensors for gloo; all ranks pass the same names / trailing shapes).
    Returns name -> tensor with leading dimension n_datasets_total, in dataset order.
    Equal shards (the usual case: datasets divisible by ranks) are gathered STRAIGHT into the result
    tensor -- no padding, no concatenation copy; unequal shards are padded to the largest one and
    trimmed afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    out = {}
    if world == 1:
        return {k: v.clone() for k, v in local.items()}
    sizes = [shard_range(n_datasets_total, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    equal = all(hi - lo == nmax for lo, hi in sizes)
    for name in sorted(local):
        t = local[name].contiguous()
        if equal:
            res = torch.empty((n_datasets_total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(res, t, group=group)
            out[name] = res
            continue
        pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        buf = torch.empty((world * nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, pad, group=group)
        parts = [buf[r * nmax: r * nmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
        out[name] = torch.cat(parts, dim=0)
    return out


def device_summaries_as_torch(batch, util_w=None):
    """Zero-copy torch views of a Batch's device summaries (for the NCCL gather)."""
    import torch

    dev = torch.device("cuda", batch.ctx.device)
    return {k: torch.as_tensor(v, device=dev) for k, v in batch.summaries_device(util_w).items()}


# ---- one process, several devices: the in-library entry (include/tppmx.h, tppmx_multi_*) -----------
def run_study_multi(model, dims, shared, datasets, chains_per_dataset: int, init_labels, init_nclu, initbeta,
                    seed: int, n_iter: int, burn: int = 0, thin: int = 1, util_w=None, device_ids=None,
                    psm: bool = True, pred: bool = True) -> dict:
    """tppmx_multi_run_study: replicate datasets (with all their chains) sharded over `device_ids`
    (None = device 0 only) inside ONE process, summaries gathered with NCCL inside the library.
    Returns psm [nd, T, ntmax, ntmax], pi_mean [nd, npred, D, T], utility [nd, T, npred], best_arm
    [nd, npred] and the timing / device info of the call."""
    import ctypes as C

    from . import _lib
    from ._abi import MAXD, Dataset, Dims, Model, Shared, c_dp, c_ip

    class Study(C.Structure):
        _fields_ = [("model", C.POINTER(Model)), ("dims", C.POINTER(Dims)), ("shared", C.POINTER(Shared)),
                    ("n_datasets", C.c_int32), ("datasets", C.POINTER(Dataset)), ("chains_per_dataset", C.c_int32),
                    ("init_labels", c_ip), ("init_nclu", c_ip), ("initbeta", c_dp), ("seed", C.c_uint64),
                    ("n_iter", C.c_int32), ("burn", C.c_int32), ("thin", C.c_int32), ("do_psm", C.c_int32),
                    ("do_pred", C.c_int32), ("util_w", c_dp)]

    class StudyOut(C.Structure):
        _fields_ = [("psm", c_dp), ("pi_mean", c_dp), ("utility", c_dp), ("best_arm", c_ip), ("ms_run", C.c_double),
                    ("ms_gather", C.c_double), ("n_devices", C.c_int32), ("used_nccl", C.c_int32)]

    L = _lib.load()
    ids = np.ascontiguousarray([0] if device_ids is None else list(device_ids), dtype=np.int32)
    h = C.c_void_p()
    rc = L.tppmx_multi_create(C.byref(h), ids.ctypes.data_as(c_ip), len(ids))
    if rc != 0:
        raise _lib.TppmxError(rc, f"tppmx_multi_create failed for devices {ids.tolist()}")
    try:
        nd, T, nt, D, npred = len(datasets), dims.T, dims.ntmax, dims.D, dims.npred
        il = np.asarray(init_labels, dtype=np.int32)
        if il.ndim == 2:
            il = np.broadcast_to(il, (nd, T, nt))
        il = np.ascontiguousarray(np.stack([np.asfortranarray(a).ravel(order="F") for a in il]))
        ic = np.asarray(init_nclu, dtype=np.int32)
        ic = np.ascontiguousarray(np.broadcast_to(ic, (nd, T)) if ic.ndim == 1 else ic)
        ib = np.ascontiguousarray(initbeta, dtype=np.float64)
        uw = np.ascontiguousarray(np.concatenate([np.asarray(util_w if util_w is not None else [], float), np.zeros(MAXD)])[:MAXD])
        arr = (Dataset * nd)(*[d.cstruct() for d in datasets])
        sh = shared.cstruct()
        st = Study()
        st.model, st.dims, st.shared = C.pointer(model), C.pointer(dims), C.pointer(sh)
        st.n_datasets, st.datasets, st.chains_per_dataset = nd, arr, chains_per_dataset
        st.init_labels, st.init_nclu, st.initbeta = il.ctypes.data_as(c_ip), ic.ctypes.data_as(c_ip), ib.ctypes.data_as(c_dp)
        st.seed, st.n_iter, st.burn, st.thin, st.do_psm, st.do_pred = seed, n_iter, burn, thin, int(psm), int(pred and npred > 0)
        st.util_w = uw.ctypes.data_as(c_dp)
        o = StudyOut()
        res = {}
        if psm:
            res["psm"] = np.zeros((nd, T, nt, nt))
            o.psm = res["psm"].ctypes.data_as(c_dp)
        if pred and npred > 0:
            pim, res["utility"], res["best_arm"] = np.zeros((nd, T, D, npred)), np.zeros((nd, T, npred)), np.zeros((nd, npred), np.int32)
            o.pi_mean, o.utility, o.best_arm = pim.ctypes.data_as(c_dp), res["utility"].ctypes.data_as(c_dp), res["best_arm"].ctypes.data_as(c_ip)
        rc = L.tppmx_multi_run_study(h, C.byref(st), C.byref(o))
        if rc != 0:
            raise _lib.TppmxError(rc, L.tppmx_multi_last_error(h).decode())
        if pred and npred > 0:
            res["pi_mean"] = pim.transpose(0, 3, 2, 1)
        res["info"] = dict(ms_run=o.ms_run, ms_gather=o.ms_gather, n_devices=o.n_devices, used_nccl=bool(o.used_nccl))
        return res
    finally:
        L.tppmx_multi_destroy(h)
