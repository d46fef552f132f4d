#!/bin/bash
# 400 k env steps of every shipped algorithm overlay on Foraging-8x8-2p-3f-v3 (64 envs, the yaml defaults otherwise): results.csv of each run
mkdir -p gpurun_out/learning
for alg in idqn vdn qmix ia2c ippo maa2c mappo; do
  timeout 100 python -m codebase_b200.run +algorithm=$alg env.name=lbforaging:Foraging-8x8-2p-3f-v3 env.time_limit=25 seed=1 \
      algorithm.total_steps=400000 algorithm.eval_interval=50000 run_dir=/tmp/learn_$alg > gpurun_out/learning/$alg.log 2>&1
  cp /tmp/learn_$alg/results.csv gpurun_out/learning/$alg.csv 2>/dev/null
  echo "$alg: $(tail -n 1 gpurun_out/learning/$alg.csv | cut -c1-80)"
done
