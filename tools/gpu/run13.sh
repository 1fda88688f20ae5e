#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/gpu/config_sweep.py --iters 8 --out gpurun_out/config_sweep_r1.json 2>&1 | tail -12 | cut -c1-600
