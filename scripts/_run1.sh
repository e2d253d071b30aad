set -x
python scripts/single_chain.py > gpurun_out/r2o_single_chain.log 2>&1; cat gpurun_out/r2o_single_chain.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 120 --csv --log-file gpurun_out/r2o_launches_dp_T3.csv python scripts/single_chain.py dp_T3 100 > gpurun_out/ncu_r2o.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 120 --csv --log-file gpurun_out/r2o_launches_default_ngg.csv python scripts/single_chain.py default_ngg 100 > gpurun_out/ncu_r2o2.log 2>&1
python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -2 gpurun_out/r2o_bench.err
