timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert |FAILED|not plausible" | head -n 10
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s19_bench.json 2> gpurun_out/s19_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/s19_bench.json")); r=d["roofline"]
    print(round(d["value"]), "ms/step", round(d["ms_per_step"],1), {k: round(v["launch_us"],1) for k,v in r["kernels"].items()}, "pass", round(r["training_pass"]["launch_us"],1), "update", round(r["whole_update"]["us_upper_bound"],1))
except Exception as e: print("ERR", e)
PY
