for f in test_a2c_gpu test_bench_gpu test_dqn_driver_gpu; do
  echo "== $f + test_dqn_gpu[257]"; timeout 600 python -m pytest tests/$f.py tests/test_dqn_gpu.py -q -x -k "not golden and not forward" 2>&1 | tail -n 3
done
echo "== driver[vdn] only + 257"; timeout 600 python -m pytest "tests/test_dqn_driver_gpu.py" tests/test_dqn_gpu.py -q -x -k "vdn or 257" 2>&1 | tail -n 3
