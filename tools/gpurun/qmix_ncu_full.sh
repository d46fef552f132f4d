#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qmix_mix|qmix_wgrad" -s 20 -c 2 -o gpurun_out/qmix_full -f python tools/qmix_time.py > gpurun_out/qmix_ncu.log 2>&1
tail -n 5 gpurun_out/qmix_ncu.log
ls -la gpurun_out/qmix_full.ncu-rep
