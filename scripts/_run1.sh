python -m pytest tests/test_gpu_fast_sweep.py tests/test_gpu_fast_probs.py -m gpu -q -x -k "double or dd_ or ineligible" 2>&1 | tail -15
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dd_ or double" 2>&1 | tail -4
python scripts/run_generic_case.py double_dipper dd_dp_T3 2>&1 | tail -4
python -c "
import torch
from treatppmx_b200 import multigpu
print(multigpu.pin_to_gpu_numa_node(0))
import os; print(sorted(os.sched_getaffinity(0))[:40])
"
