#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/gpu/config_sweep.py --configs C2,C3,C4,C5 --iters 16 --topologies reference,ploc --modes batch --out gpurun_out/config_sweep_r1_final.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print(r['config'], r['topology'], '%.3f ms/it' % r['ms_per_iteration'], '%.1f Mpaths/s' % r['mpaths_per_s'], '%.1f Mrays/s' % r['mrays_per_s'], 'topo build %.0f ms' % r['topology_build_ms'])
"
