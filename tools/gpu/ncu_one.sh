#!/bin/bash
TAG=$1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 18 -c 1 -o gpurun_out/ncu_${TAG}_trace1 -f python tools/gpu/wf_sweep.py --iters 1 --only-wavefront > gpurun_out/ncu_${TAG}_trace1.log 2>&1
ls -la gpurun_out/*.ncu-rep
