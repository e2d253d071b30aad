for r in 1 2 3 4 5 6 7 8; do python scripts/single_chain.py dp_T3 2>&1 | grep -E "1100 iterations" | cut -c45-100 | tr '\n' ' '; echo; done
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_shim.py tests/test_reference_golden.py tests/test_simupats.py -m gpu -q 2>&1 | tail -2
