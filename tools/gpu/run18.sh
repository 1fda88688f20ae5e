#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/gpu/config_sweep.py --configs C2,C5 --iters 16 --topologies reference,ploc --modes batch --out gpurun_out/topology_sweep.json 2>&1 | cut -c1-700
