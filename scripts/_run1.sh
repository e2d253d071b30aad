python -m pytest tests/test_gpu_fast_sweep.py tests/test_gpu_fast_probs.py -m gpu -q -x -k "nnig or ineligible" 2>&1 | tail -15
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nnig" 2>&1 | tail -8
python scripts/run_generic_case.py nnig nnig_ngg_T3 2>&1 | tail -4
