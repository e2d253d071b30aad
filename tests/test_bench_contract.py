"""bench.py's reference arm (the CPU legs: oracle port of the sweep phase, and the reference's own
dm_ppmx_ct where oracle/_ref exists) prints ONE JSON line with the contract's keys.  Runs on CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--datasets", "2", "--chains-per-dataset", "1", "--cpu-scale", "0.02"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "patient_reallocations_per_sec" and d["unit"] == "reallocations/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config"):
        assert k in d, k
    assert d["value"] > 0 and d["higher_is_better"] is True and d["dtype"] == "f64" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    # cores = the threads actually used: affinity mask capped by the cgroup quota, not os.cpu_count()
    assert cb["host"]["threads_used"] == cb["cores"] <= cb["host"]["affinity"] <= cb["host"]["cpu_count"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    fi = d["full_iteration"]
    assert fi["cpu_port_gibbs_iterations_per_sec"] > 0
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref.so")):
        assert fi["reference_kind"] == "reference" and fi["reference_gibbs_iterations_per_sec"] > 0


def test_profile_records_are_keyed_by_shape():
    """roofline.traffic / issue.* come from an ncu capture of the SAME shape or are null."""
    sys.path.insert(0, ROOT)
    import bench
    from treatppmx_b200._abi import Dims
    d = Dims()
    d.nobs, d.T, d.P = 152, 3, 10
    rec = bench.profile_record("sweep", bench.shape_key(d, 1024, 20))
    assert rec and rec["dram_bytes_per_launch"] > 0 and rec["warp_instructions_per_launch"] > 0 and "ncu" in rec["source"]
    d.nobs, d.P = 20000, 50
    assert bench.profile_record("sweep", bench.shape_key(d, 1024, 20)) is None
