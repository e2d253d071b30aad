python -m pytest tests/test_gpu_fast_sweep.py tests/test_gpu_fast_probs.py -m gpu -q -x  2>&1 | tail -15

