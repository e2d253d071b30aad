python -m pytest tests/test_gpu_fast_sweep.py tests/test_gpu_fast_probs.py -m gpu -q -x -k "cat" 2>&1 | tail -15
python -m pytest tests/test_gpu_parity.py tests/test_gpu_update.py -m gpu -q -k "cat" 2>&1 | tail -4
python scripts/run_generic_case.py categorical categorical_T3 cat_calib2 2>&1 | tail -4
