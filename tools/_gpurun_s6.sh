timeout 300 python tools/debug_vdn257.py > gpurun_out/s6_dbg.txt 2>&1; tail -n 12 gpurun_out/s6_dbg.txt
timeout 600 python -m pytest tests/test_tc_backward_gpu.py tests/test_dqn_gpu.py -q -x 2>&1 | tail -n 8
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/s6_bench.json",):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, round(d["value"]), "ms/step", round(d["ms_per_step"],1), {k: round(v["launch_us"],1) for k,v in r["kernels"].items()}, "pass", round(r["training_pass"]["launch_us"],1), "update", round(r["whole_update"]["us_upper_bound"],1))
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 gpurun_out/s6_bench.err
MARL_B200_SO=$PWD/codebase_b200/csrc/libmarlb200_ts.so timeout 300 python tools/ts_timeline.py > gpurun_out/s6_timeline.txt 2>&1
