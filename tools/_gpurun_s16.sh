timeout 600 python tools/learning_curve.py --arm gpu --seeds 4 --envs 64 --total 1200000 --eval-every 100000 --out gpurun_out/r2_learning_curve_gpu_1p2M.json > /dev/null 2> gpurun_out/r2_lc.err; tail -n 2 gpurun_out/r2_lc.err
timeout 400 python bench.py --config ia2c --steps 5 --warmup 3 > gpurun_out/r2_bench_ia2c.json 2> gpurun_out/r2_bench_ia2c.err; cut -c1-300 gpurun_out/r2_bench_ia2c.json
timeout 400 python bench.py --config vdn15 --steps 3 --warmup 3 > gpurun_out/r2_bench_vdn15.json 2> gpurun_out/r2_bench_vdn15.err; cut -c1-300 gpurun_out/r2_bench_vdn15.json
