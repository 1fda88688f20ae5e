#!/bin/bash
# ncu --set full captures of the wavefront kernels inside one PT iteration of the bench workload.
# usage: tools/gpu/ncu_wf.sh <tag>    -> gpurun_out/ncu_<tag>_{trace1,shade0,trace0}.ncu-rep
set -x
TAG=${1:-wf}
mkdir -p gpurun_out
# launches of k_wf_*: per iteration generate, (trace, shade) x16, trace = 34; skip the warm-up iteration (34) and pick inside the next one
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 18 -c 1 -o gpurun_out/ncu_${TAG}_trace1 -f python tools/gpu/wf_sweep.py --iters 1 --only-wavefront > gpurun_out/ncu_${TAG}_trace1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 17 -c 1 -o gpurun_out/ncu_${TAG}_trace0 -f python tools/gpu/wf_sweep.py --iters 1 --only-wavefront > gpurun_out/ncu_${TAG}_trace0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_shade -s 16 -c 1 -o gpurun_out/ncu_${TAG}_shade0 -f python tools/gpu/wf_sweep.py --iters 1 --only-wavefront > gpurun_out/ncu_${TAG}_shade0.log 2>&1
ls -la gpurun_out/*.ncu-rep
