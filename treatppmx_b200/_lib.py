"""Loader of the CUDA library (treatppmx_b200/libtppmx.so, built by csrc/build.sh).
There is no CPU fallback: if the library is missing this raises, loudly."""
from __future__ import annotations

import ctypes as C
import os

from ._abi import Dataset, Dims, Model, Shared, State, c_dp, c_ip

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("TPPMX_LIB") or os.path.join(_HERE, "libtppmx.so")  # TPPMX_LIB: e.g. the -DTPPMX_CHECK build

# every symbol include/tppmx.h declares
SYMBOLS = [
    "tppmx_version", "tppmx_create", "tppmx_destroy", "tppmx_last_error", "tppmx_probe_math",
    "tppmx_batch_create", "tppmx_batch_destroy", "tppmx_batch_init",
    "tppmx_batch_set_state", "tppmx_batch_get_state", "tppmx_batch_get_partitions",
    "tppmx_score_patient",
    "tppmx_batch_sweep", "tppmx_batch_sweep_probe", "tppmx_batch_set_option", "tppmx_batch_get_info",
    "tppmx_batch_last_timing", "tppmx_batch_predict",
    "tppmx_batch_accumulate", "tppmx_batch_summaries", "tppmx_batch_summaries_device",
    "tppmx_batch_reset_summaries", "tppmx_batch_update", "tppmx_batch_run", "tppmx_batch_get_acceptance",
    "tppmx_dm_ppmx_ct", "tppmx_prior_ppmx_core", "tppmx_batch_npc", "tppmx_set_interrupt_flag",
    "tppmx_batch_point_estimate",
    "tppmx_multi_create", "tppmx_multi_destroy", "tppmx_multi_device_count", "tppmx_multi_last_error",
    "tppmx_multi_run_study",
]

_lib = None


class TppmxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tppmx error {code}: {msg}")
        self.code = code


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} not found: the CUDA library is the only compute path of treatppmx_b200 "
            "(no CPU fallback). Build it with treatppmx_b200/csrc/build.sh or "
            "python -c 'import __graft_entry__ as g; g.build()'.")
    L = C.CDLL(SO_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    L.tppmx_version.restype = C.c_int
    L.tppmx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.tppmx_destroy.argtypes = [vp]
    L.tppmx_last_error.argtypes = [vp]
    L.tppmx_last_error.restype = C.c_char_p
    L.tppmx_probe_math.argtypes = [vp, i32, c_dp, i32, c_dp]
    L.tppmx_batch_create.argtypes = [vp, C.POINTER(Model), C.POINTER(Dims), C.POINTER(Shared), i32,
                                     C.POINTER(Dataset), i32, c_ip, c_ip, C.POINTER(vp)]
    L.tppmx_batch_destroy.argtypes = [vp]
    L.tppmx_batch_init.argtypes = [vp, u64, c_ip, c_ip, c_dp]
    L.tppmx_batch_set_state.argtypes = [vp, i32, C.POINTER(State)]
    L.tppmx_batch_get_state.argtypes = [vp, i32, C.POINTER(State)]
    L.tppmx_batch_get_partitions.argtypes = [vp, c_ip, c_ip]
    L.tppmx_score_patient.argtypes = [vp, i32, i32, i32, u64, i32, c_dp, c_dp, c_ip]
    L.tppmx_batch_sweep.argtypes = [vp, u64, i32, i32]
    L.tppmx_batch_sweep_probe.argtypes = [vp, u64, i32, i32, i32, c_dp, c_ip]
    L.tppmx_batch_set_option.argtypes = [vp, i32, i32]
    L.tppmx_batch_get_info.argtypes = [vp, i32, c_ip]
    L.tppmx_batch_last_timing.argtypes = [vp, c_dp, c_ip]
    L.tppmx_batch_predict.argtypes = [vp, u64, i32, c_dp, c_dp, c_ip]
    L.tppmx_batch_accumulate.argtypes = [vp, u64, i32, i32, i32]
    L.tppmx_batch_summaries.argtypes = [vp, c_dp, c_dp, c_dp, c_dp, c_ip]
    L.tppmx_batch_summaries_device.argtypes = [vp, c_dp, C.POINTER(c_dp), C.POINTER(c_dp), C.POINTER(c_dp),
                                               C.POINTER(c_ip)]
    L.tppmx_batch_reset_summaries.argtypes = [vp]
    L.tppmx_batch_update.argtypes = [vp, u64, i32]
    L.tppmx_batch_run.argtypes = [vp, u64, i32, i32, i32, i32, i32, i32]
    L.tppmx_batch_get_acceptance.argtypes = [vp, i32, c_ip, c_ip]
    L.tppmx_dm_ppmx_ct.argtypes = [vp, vp, vp]
    L.tppmx_prior_ppmx_core.argtypes = [vp, vp, vp]
    L.tppmx_batch_point_estimate.argtypes = [vp, i32, i32, i32, c_ip, c_ip, c_dp]
    L.tppmx_multi_create.argtypes = [C.POINTER(vp), c_ip, i32]
    L.tppmx_multi_destroy.argtypes = [vp]
    L.tppmx_multi_device_count.argtypes = [vp]
    L.tppmx_multi_last_error.argtypes = [vp]
    L.tppmx_multi_last_error.restype = C.c_char_p
    L.tppmx_multi_run_study.argtypes = [vp, vp, vp]
    L.tppmx_batch_npc.argtypes = [vp, c_ip, c_ip, c_ip]
    L.tppmx_set_interrupt_flag.argtypes = [vp, c_ip, i32]
    _lib = L
    return L
