#!/bin/bash
# wavefront LPT / BPT: GPU tests, then the scheduler comparison of the three renderers on the Sponza-shaped 1080p workload
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
(time python -m pytest tests -m gpu -x -q) 2>&1 | tail -15
timeout 600 python tools/gpu/renderer_sweep.py --iters 12 --out gpurun_out/renderer_sweep_r1f.json 2>&1 | cut -c1-420 | tail -40
