set -x
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2b_tests.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -15 gpurun_out/r2b_tests.log; cut -c1-300 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
