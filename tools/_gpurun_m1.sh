timeout 600 python -m pytest tests/test_peer_exchange_gpu.py -q 2>&1 | grep -E "passed|failed|Error|assert |FAILED|skipped" | head -n 5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/m1_bench_2gpu.json 2> gpurun_out/m1_bench_2gpu.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/m1_bench_2gpu.json"))
    print(round(d["value"]), d["n_gpus"], round(d["ms_per_step"],1), d.get("ranks_bit_identical"), {k:v for k,v in d.get("multi_gpu",{}).items() if "peer" in k or "identical" in k})
except Exception as e: print("ERR", e)
PY
tail -n 3 gpurun_out/m1_bench_2gpu.err
