#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qmix.py "tests/test_dqn_driver_gpu.py::test_driver_writes_reference_schema" tests/test_ppo.py -m gpu -x -q > gpurun_out/qmix_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/qmix_tests.log
timeout 120 python tools/qmix_time.py > gpurun_out/qmix_time.log 2>&1
tail -n 15 gpurun_out/qmix_tests.log; cat gpurun_out/qmix_time.log
