WATCHDOG_S=25 timeout -s KILL 90 python tools/debug_hang.py 2>&1 | cut -c1-250 | tail -n 6
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/s9_bench.json",):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, round(d["value"]), "ms/step", round(d["ms_per_step"],1), {k: round(v["launch_us"],1) for k,v in r["kernels"].items()}, "pass", round(r["training_pass"]["launch_us"],1), "update", round(r["whole_update"]["us_upper_bound"],1))
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 gpurun_out/s9_bench.err
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 6
