set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"tc_|reduce_adam" -s 1100 -c 60 --csv --log-file gpurun_out/r2a_traffic_warm.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --updates-per-iter 64 > gpurun_out/r2a_ncu_bench.json 2> gpurun_out/r2a_ncu.err
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_bench.json | cut -c1-600
