# One gpurun call at the end of a round: tests, smoke, both bench arms, launch lists and the ncu capture of the dominant
# kernel -> gpurun_out/r2z_* (copied / summarised into profiles/ afterwards).  bash scripts/gpu_round_capture.sh
set -x
R=r2z
python -m pytest tests -m gpu -q > gpurun_out/${R}_gputests.log 2>&1; tail -3 gpurun_out/${R}_gputests.log
TPPMX_LIB=$PWD/treatppmx_b200/libtppmx_check.so python -m pytest tests/test_gpu_fast_sweep.py tests/test_gpu_fast_probs.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_large_n.py tests/test_gpu_update.py -m gpu -q > gpurun_out/${R}_gputests_check.log 2>&1; tail -2 gpurun_out/${R}_gputests_check.log
python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -2 gpurun_out/${R}_bench.err
if [ "${FULL:-0}" = "1" ]; then
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-configs"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches.csv $B > gpurun_out/ncu1.log 2>&1
$B > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep_fast -s 52 -c 1 -o gpurun_out/${R}_sweep_fast_timed $B > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
python scripts/run_iters.py 5 > gpurun_out/plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 121 -c 60 --csv --log-file gpurun_out/${R}_launches_run.csv python scripts/run_iters.py 5 > gpurun_out/ncu3.log 2>&1
fi
python scripts/run_iters.py 40 | tail -1
python scripts/single_chain.py > gpurun_out/${R}_single_chain.log 2>&1; cat gpurun_out/${R}_single_chain.log
