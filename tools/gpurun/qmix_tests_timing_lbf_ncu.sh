#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qmix.py "tests/test_dqn_driver_gpu.py::test_driver_writes_reference_schema" -m gpu -x -q > gpurun_out/qmix_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/qmix_tests.log
MARL_QMIX_WGRAD_TILES=1 timeout 300 python -m pytest tests/test_qmix.py -m gpu -x -q > gpurun_out/qmix_tests_tiles.log 2>&1
echo "tests (tile form) exit $?" >> gpurun_out/qmix_tests_tiles.log
timeout 120 python tools/qmix_time.py > gpurun_out/qmix_time.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qmix -s 40 -c 12 --csv --log-file gpurun_out/qmix_launches.csv python tools/qmix_time.py > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lbf_step -s 130 -c 1 -o gpurun_out/lbf_step_2p20 -f python tools/bench_extra.py env_sweep > gpurun_out/lbf_sweep_under_ncu.log 2>&1
tail -n 4 gpurun_out/qmix_tests.log; tail -n 4 gpurun_out/qmix_tests_tiles.log; cat gpurun_out/qmix_time.log; cut -d, -f5,12- gpurun_out/qmix_launches.csv | tail -n 8; tail -n 3 gpurun_out/lbf_sweep_under_ncu.log
