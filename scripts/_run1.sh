python -m pytest tests/test_gpu_update.py tests/test_gpu_summaries.py tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_shim.py tests/test_gpu_edge_cases.py tests/test_gpu_large_n.py tests/test_reference_golden.py -m gpu -q -x 2>&1 | tail -3
python scripts/run_iters.py 40 28
python scripts/run_iters.py 10 10000
