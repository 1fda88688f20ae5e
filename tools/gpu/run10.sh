#!/bin/bash
# source-level capture of the steady-state trace wave of a batch (wave 58 of a 24-iteration batch carries 16 iterations)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 58 -c 1 -o gpurun_out/ncu_r1g_trace -f python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1g_trace.log 2>&1
ls -la gpurun_out/
