"""Generates tests/golden/ref_<case>.npz: the 18-element return list of the REFERENCE's own
dm_ppmx_ct (src/dm_ppmx_ct.cpp:10-1471, compiled unmodified into oracle/_ref/libref.so against
oracle/refshim/) for a handful of parity cases, with the Philox-contract variates (recorded from the
oracle, tests/refrun.py:Tape) in place of R's RNG.  Needs /root/reference; run from the repo root:

    python tests/golden/make_reference_golden.py

The files are what the oracle (CPU test) and the CUDA drop-in tppmx_dm_ppmx_ct (GPU test) must
reproduce where the reference itself is not available (tests/test_reference_golden.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

import refrun  # noqa: E402

REF_GOLDEN = {  # name -> (iters, burn, thin, seed)
    "default_ngg": (36, 12, 3, 99),
    "dp_T3": (36, 12, 3, 99),
    "nnig": (30, 10, 2, 5),
    "calib1": (30, 10, 2, 6),
    "calib2_sqrt": (30, 10, 2, 7),
    "noreuse": (30, 10, 2, 8),
    "categorical": (30, 10, 2, 9),
    "Q3_D3": (30, 10, 2, 10),
    # the shapes the fast kernel's special instantiations run at (n = 152, T = 3, P = 10)
    "calib1_dp_T3": (30, 10, 2, 11),
    "nnig_ngg_T3": (30, 10, 2, 12),
    "categorical_T3": (30, 10, 2, 13),
    "dd_dp_T3": (30, 10, 2, 14),
    "noreuse_T3": (30, 10, 2, 15),
    # BASELINE config 1: the bundled simupats data (tests/golden/simupats_config1.npz), ppmxct()'s default
    # chain length iter = 1100 / burn = 100, thinned to 50 saved iterations to keep the file small
    "simupats_ngg": (1100, 100, 20, 121),
    "simupats_dp": (1100, 100, 20, 122),
}


def path(name):
    return os.path.join(HERE, f"ref_{name}.npz")


def main():
    R = refrun.RefLib()
    only = sys.argv[1:]
    for name, (iters, burn, thin, seed) in REF_GOLDEN.items():
        if only and name not in only:
            continue
        args, _, case = refrun.reference_args(name, iters, burn, thin, True)
        with refrun.Tape() as tape:
            refrun.oracle_run(case, iters, burn, thin, seed)
        got = R.dm_ppmx_ct(args, tape)
        assert got.pop("tape_used") == len(tape)
        got["WAIC"] = np.array(got["WAIC"])
        np.savez_compressed(path(name), **got)
        print(name, "draws", len(tape), os.path.getsize(path(name)), "bytes")


if __name__ == "__main__":
    main()
