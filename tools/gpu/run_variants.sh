#!/bin/bash
# GPU box: times every gpugi_b200/variants/libgpugi_b200_<tag>.so on the batch schedule (and checks trace parity once per variant).
# env: RENDERERS (default PT), ITERS (24), TUNES = "refill:minInner ..." (default "16:20"), REPS (2)
mkdir -p gpurun_out
cp gpugi_b200/libgpugi_b200.so /tmp/lib_default.so
for lib in gpugi_b200/variants/libgpugi_b200_*.so; do
  tag=$(basename $lib .so | sed 's/libgpugi_b200_//')
  cp $lib gpugi_b200/libgpugi_b200.so
  par=$(python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1)
  for tune in ${TUNES:-16:20}; do
  for rep in $(seq 1 ${REPS:-2}); do
  GPUGI_WF_REFILL=${tune%%:*} GPUGI_WF_MIN_INNER=${tune##*:} timeout 300 python tools/gpu/renderer_sweep.py --iters ${ITERS:-24} --renderers ${RENDERERS:-PT} --out gpurun_out/variant_$tag.json 2>&1 | grep '"batch"' | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    print('$tag tune $tune %-4s %-10s %-10s %7.3f ms/it  %7.1f Mpaths/s | $par' % (r['renderer'], r['scheduler'], r['mode'], r['ms_per_iteration'], r['mpaths_per_s']))
"
  done
  done
done
cp /tmp/lib_default.so gpugi_b200/libgpugi_b200.so
