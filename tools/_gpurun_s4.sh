set -x
timeout 600 python -m pytest tests/test_dqn_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/s4_dqn_alone.log
timeout 600 python -m pytest tests/test_dqn_gpu.py -q -x -k "random_batches" 2>&1 | tail -5 > gpurun_out/s4_dqn_rb.log
timeout 600 python -m pytest tests/test_dqn_gpu.py -q -x -k "1-False-257-2" 2>&1 | tail -5 > gpurun_out/s4_dqn_one.log
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active -k regex:"tc_|reduce_adam" -s 1100 -c 40 --csv --log-file gpurun_out/s4_traffic_warm.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --updates-per-iter 64 > gpurun_out/s4_ncu_bench.json 2> gpurun_out/s4_ncu.err
tail -3 gpurun_out/s4_dqn_alone.log gpurun_out/s4_dqn_rb.log gpurun_out/s4_dqn_one.log
