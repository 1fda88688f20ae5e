#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu/config_sweep.py --configs C2,C5 --iters 16 --topologies ploc --radii 4,8,32,64 --modes batch --out gpurun_out/topology_radius_sweep.json 2>&1 | cut -c1-420
timeout 900 python tools/gpu/config_sweep.py --configs C3,C4 --iters 8 --topologies reference,ploc --modes batch --out gpurun_out/topology_sweep_c3_c4.json 2>&1 | cut -c1-420
