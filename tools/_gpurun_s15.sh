# launch list of the default bench (cold-cache, serialised) + one --set full capture of each on-chip kernel inside the running pipeline
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --updates-per-iter 64 > gpurun_out/r2_launches_bench.json 2> gpurun_out/r2_launches.err
timeout 900 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"tc_dqn_fwd3|tc_dh1w1|tc_dw2" -s 150 -c 3 -o gpurun_out/r2_onchip_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --updates-per-iter 64 > gpurun_out/r2_full_bench.json 2> gpurun_out/r2_full.err
ls -la gpurun_out/ | tail -n 8
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; cut -c1-400 gpurun_out/r2_bench_default.json
