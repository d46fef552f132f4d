set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/s2_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
MARL_B200_SO=$PWD/codebase_b200/csrc/libmarlb200_dw2.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s2_bench_dw2.json 2> gpurun_out/s2_bench_dw2.err
tail -8 gpurun_out/s2_tests.log
python - <<'PY'
import json
for f in ("gpurun_out/s2_bench.json","gpurun_out/s2_bench_dw2.json"):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, round(d["value"]), "ms/step", round(d["ms_per_step"],1), {k: round(v["launch_us"],1) for k,v in r["kernels"].items()}, "pass", round(r["training_pass"]["launch_us"],1), "update", round(r["whole_update"]["us_upper_bound"],1))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/s2_bench.err
