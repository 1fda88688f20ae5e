#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:k_wf_trace -s 40 -c 1 -o gpurun_out/ncu_r1_c5_trace -f python tools/gpu/config_sweep.py --configs C5 --iters 20 > gpurun_out/ncu_r1_c5.log 2>&1
tail -3 gpurun_out/ncu_r1_c5.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep | tail -3
