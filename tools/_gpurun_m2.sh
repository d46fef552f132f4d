timeout 300 python -m pytest tests/test_peer_exchange_gpu.py -q 2>&1 | grep -E "passed|failed|Error|assert |FAILED|skipped" | head -n 5
for se in 1 0; do
MARL_SPLIT_EXCHANGE=$se timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$se bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/m2_bench_2gpu_se$se.json 2> gpurun_out/m2_bench_2gpu_se$se.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/m2_bench_2gpu_se$se.json"))
    print("split=$se", round(d["value"]), d["n_gpus"], round(d["ms_per_step"],1), d.get("ranks_bit_identical"), {k:v for k,v in d.get("multi_gpu",{}).items() if "peer" in k})
except Exception as e: print("ERR", e)
PY
done
tail -n 3 gpurun_out/m2_bench_2gpu_se1.err
