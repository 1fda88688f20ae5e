#!/bin/bash
# A/B of library variants on the other configs: run23.sh "C3,C5" tagA tagB ...
mkdir -p gpurun_out
CFG=$1; shift
cp gpugi_b200/libgpugi_b200.so /tmp/lib_default.so
for tag in "$@"; do
  cp gpugi_b200/variants/libgpugi_b200_$tag.so gpugi_b200/libgpugi_b200.so
  timeout 600 python tools/gpu/config_sweep.py --configs $CFG --iters 12 --topologies ploc --modes batch --out gpurun_out/cfg_variant_$tag.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print('$tag', r['config'], '%.3f ms/it' % r['ms_per_iteration'], '%.1f Mpaths/s' % r['mpaths_per_s'], 'mean', r['mean_radiance'])
"
done
cp /tmp/lib_default.so gpugi_b200/libgpugi_b200.so
