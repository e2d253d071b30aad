"""treatppmx_b200 — B200-native (sm_100a) PPMx reallocation sweep + posterior predictive.

The compute path is the CUDA library behind include/tppmx.h (treatppmx_b200/libtppmx.so);
this package is the thin host-side mirror of the reference's interface for that path.
There is no CPU fallback: using the engine without the built library raises.
"""
from ._abi import (Dims, HostDataset, HostShared, HostState, Model, default_model)  # noqa: F401

__version__ = "0.1.0"
