#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/final_tests.log
timeout 200 python tools/bench_extra.py env_sweep > gpurun_out/lbf_sweep.jsonl 2> gpurun_out/lbf_sweep.err
timeout 300 ncu --set full --clock-control none -k regex:lbf_step -s 130 -c 1 -o gpurun_out/lbf_step_2p20_padded -f python tools/bench_extra.py env_sweep > /dev/null 2>&1
tail -n 3 gpurun_out/final_tests.log; cat gpurun_out/lbf_sweep.jsonl | cut -c1-260
