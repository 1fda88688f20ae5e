#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu/config_sweep.py --configs C2,C5 --iters 16 --topologies ploc --radii 1,2,3,6 --modes batch --out gpurun_out/topology_radius_sweep2.json 2>&1 | cut -c1-420
timeout 300 python tools/gpu/sanitize.py 2>&1 | tail -5
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer_$tool.log python tools/gpu/sanitize.py > gpurun_out/sanitizer_$tool.out 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_$tool.out | cut -c1-300; tail -4 gpurun_out/sanitizer_$tool.log | cut -c1-300
done
