timeout 600 python -m pytest tests/test_episode_lengths_gpu.py tests/test_tc_backward_gpu.py tests/test_dqn_gpu.py tests/test_wide_obs_gpu.py tests/test_dqn_driver_gpu.py -q 2>&1 | grep -E "passed|failed|Error|assert |FAILED|not plausible" | head -n 10
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s18_bench.json 2> gpurun_out/s18_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/s18_bench.json")); r=d["roofline"]
    print(round(d["value"]), "ms/step", round(d["ms_per_step"],1), {k: round(v["launch_us"],1) for k,v in r["kernels"].items()}, "pass", round(r["training_pass"]["launch_us"],1), "update", round(r["whole_update"]["us_upper_bound"],1))
except Exception as e: print("ERR", e)
PY
MARL_B200_SO=$PWD/codebase_b200/csrc/libmarlb200_ts.so timeout 300 python tools/ts_timeline.py > gpurun_out/s18_timeline.txt 2>&1
