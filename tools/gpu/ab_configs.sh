#!/bin/bash
# GPU box: A/B of the device-library variants built by tools/gpu/variants.sh on the batch schedule of the BASELINE configs.
#   CONFIGS="C3 C4" ITERS=24 REPS=3 tools/gpu/ab_configs.sh [tag ...]     (default: every gpugi_b200/variants/<tag>/ + the product build as "base")
mkdir -p gpurun_out
tags=${@:-base $(ls gpugi_b200/variants)}
for c in ${CONFIGS:-C2}; do
for tag in $tags; do
  if [ "$tag" = base ]; then unset GPUGI_B200_LIBDIR; else export GPUGI_B200_LIBDIR=$(pwd)/gpugi_b200/variants/$tag; fi
  python tools/gpu/prof_batch.py --config $c --iters ${ITERS:-24} --repeat ${REPS:-3} --kernel-timing 2>/dev/null | python -c "
import sys, json
rows = [json.loads(l) for l in sys.stdin if l.startswith('{')]
if rows:
    b = min(rows, key=lambda r: r['ms_per_iteration'])
    print('%-4s %-12s %8.3f ms/it %8.1f Mpaths/s  %s' % ('$c', '$tag', b['ms_per_iteration'], b['mpaths_per_s'], ' '.join('%s %.3f' % kv for kv in b.get('kernel_ms_per_iteration', {}).items())))
else:
    print('$c $tag FAILED')
"
done
done
unset GPUGI_B200_LIBDIR
