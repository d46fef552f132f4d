#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/final_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
timeout 120 python tools/qmix_time.py > gpurun_out/qmix_time.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qmix -s 40 -c 8 --csv --log-file gpurun_out/qmix_launches.csv python tools/qmix_time.py > /dev/null 2>&1
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -n 6 gpurun_out/final_tests.log; tail -n 2 gpurun_out/final_smoke.log; cat gpurun_out/qmix_time.log; cut -d, -f5,12- gpurun_out/qmix_launches.csv | tail -n 4; cat gpurun_out/final_bench.json | cut -c1-700
