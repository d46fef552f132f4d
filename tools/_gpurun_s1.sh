set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/s1_tests.log
MARL_TC_PINGPONG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s1_bench_pp1.json 2> gpurun_out/s1_bench_pp1.err
MARL_TC_PINGPONG=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s1_bench_pp0.json 2> gpurun_out/s1_bench_pp0.err
tail -12 gpurun_out/s1_tests.log
python - <<'PY'
import json
for f in ("gpurun_out/s1_bench_pp1.json","gpurun_out/s1_bench_pp0.json"):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, round(d["value"]), "ms/step", round(d["ms_per_step"],1), {k: round(v["launch_us"],1) for k,v in r["kernels"].items()}, "pass", round(r["training_pass"]["launch_us"],1), "update", round(r["whole_update"]["us_upper_bound"],1))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/s1_bench_pp1.err
