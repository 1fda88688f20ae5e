#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer_$tool.log python tools/gpu/sanitize.py > gpurun_out/sanitizer_$tool.out 2>&1
  echo "$tool rc=$?"; tail -1 gpurun_out/sanitizer_$tool.out | cut -c1-200; tail -2 gpurun_out/sanitizer_$tool.log | cut -c1-200
done
