#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/final_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
grep -E "passed|failed|Error|error" gpurun_out/final_tests.log | tail -n 12; tail -n 2 gpurun_out/final_smoke.log; cut -c1-200 gpurun_out/final_bench.json
