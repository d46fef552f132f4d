timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert |FAILED|not plausible" | head -n 20
timeout 200 python tools/debug_vdn257.py 2>&1 | grep "onchip=1" | cut -c1-200 | tail -n 9
for so in libmarlb200.so libmarlb200_rn.so; do
MARL_B200_SO=$PWD/codebase_b200/csrc/$so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s14_bench_$so.json 2> gpurun_out/s14_bench.err
python - <<PY
import json
f="gpurun_out/s14_bench_$so.json"
try:
    d=json.load(open(f)); r=d["roofline"]
    print(f, round(d["value"]), "ms/step", round(d["ms_per_step"],1), {k: round(v["launch_us"],1) for k,v in r["kernels"].items()}, "pass", round(r["training_pass"]["launch_us"],1), "update", round(r["whole_update"]["us_upper_bound"],1))
except Exception as e: print(f, "ERR", e)
PY
done
