#!/bin/bash
# GPU box: times every gpugi_b200/variants/<tag>/ build on the batch schedule (and checks trace parity once per variant).
# env: RENDERERS (default PT), ITERS (24), TUNES = "refill:minInner ..." (default "24:16"), REPS (2), CONFIGS (config_sweep.py --configs, default none)
mkdir -p gpurun_out
for dir in gpugi_b200/variants/*/; do
  tag=$(basename $dir)
  export GPUGI_B200_LIBDIR=$(pwd)/$dir
  par=$(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1)
  for tune in ${TUNES:-24:16}; do
  for rep in $(seq 1 ${REPS:-2}); do
  GPUGI_WF_REFILL=${tune%%:*} GPUGI_WF_MIN_INNER=${tune##*:} timeout 300 python tools/gpu/renderer_sweep.py --iters ${ITERS:-24} --renderers ${RENDERERS:-PT} --out gpurun_out/variant_$tag.json 2>&1 | grep '"batch"' | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    print('$tag tune $tune %-4s %-10s %-10s %7.3f ms/it  %7.1f Mpaths/s | $par' % (r['renderer'], r['scheduler'], r['mode'], r['ms_per_iteration'], r['mpaths_per_s']))
"
  done
  done
  if [ -n "$CONFIGS" ]; then timeout 900 python tools/gpu/config_sweep.py --configs $CONFIGS --out gpurun_out/variant_${tag}_configs.json 2>&1 | grep '^{' | sed "s/^/$tag /"; fi
done
unset GPUGI_B200_LIBDIR
