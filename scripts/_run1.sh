python -m pytest tests/test_gpu_dropin.py tests/test_gpu_shim.py tests/test_reference_golden.py tests/test_simupats.py tests/test_gpu_edge_cases.py -m gpu -q 2>&1 | tail -3
export TPPMX_TRACE=1
for r in 1 2 3; do python scripts/single_chain.py dp_T3 2>&1 | grep -E "1100 iterations" | cut -c1-135 | tr '\n' ' '; echo; done
python scripts/single_chain.py default_ngg 2>&1 | grep -E "1100 iterations" | cut -c1-135 | tr '\n' ' '; echo
TPPMX_DROPIN_NO_GRAPH=1 python scripts/single_chain.py dp_T3 2>&1 | grep -E "1100 iterations" | cut -c1-135 | tr '\n' ' '; echo
