#!/bin/bash
# quick A/B of library variants on the bench workload (PT batch) without the parity suite per variant: run25.sh tag ...
mkdir -p gpurun_out
cp gpugi_b200/libgpugi_b200.so /tmp/lib_default.so
for rep in 1 2; do
for tag in "$@"; do
  cp gpugi_b200/variants/libgpugi_b200_$tag.so gpugi_b200/libgpugi_b200.so
  timeout 300 python tools/gpu/config_sweep.py --configs C2 --iters 24 --topologies ploc --modes batch --out gpurun_out/cfg_variant_$tag.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print('$tag', r['config'], '%.3f ms/it' % r['ms_per_iteration'], '%.1f Mpaths/s' % r['mpaths_per_s'], 'mean', r['mean_radiance'])
"
done
done
cp /tmp/lib_default.so gpugi_b200/libgpugi_b200.so
