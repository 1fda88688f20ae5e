#!/bin/bash
# GPU box: times every gpugi_b200/variants/libgpugi_b200_<tag>.so on the PT batch schedule (and checks trace parity once per variant)
mkdir -p gpurun_out
cp gpugi_b200/libgpugi_b200.so /tmp/lib_default.so
for lib in gpugi_b200/variants/libgpugi_b200_*.so; do
  tag=$(basename $lib .so | sed 's/libgpugi_b200_//')
  cp $lib gpugi_b200/libgpugi_b200.so
  par=$(python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1)
  for rep in 1 2; do
  timeout 300 python tools/gpu/renderer_sweep.py --iters ${ITERS:-24} --renderers ${RENDERERS:-PT} --out gpurun_out/variant_$tag.json 2>&1 | grep '"batch"' | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    print('$tag %-4s %-10s %-10s %7.3f ms/it  %7.1f Mpaths/s | $par' % (r['renderer'], r['scheduler'], r['mode'], r['ms_per_iteration'], r['mpaths_per_s']))
"
  done
done
cp /tmp/lib_default.so gpugi_b200/libgpugi_b200.so
