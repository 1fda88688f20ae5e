#!/bin/bash
# Builds A/B variants of the device library (CPU box; nvcc cross-compiles): tools/gpu/variants.sh tag "-DFLAG=0 ..." [tag flags ...]
# -> gpugi_b200/variants/<tag>/{libgpugi_b200.so, libgpugi_b200_host.so} (git-ignored, travels with gpurun).  A variant is selected with
# GPUGI_B200_LIBDIR=gpugi_b200/variants/<tag> (gpugi_b200/_capi.py); the product libraries are never overwritten.
# tools/gpu/run_variants.sh times them on the GPU box.
set -e
cd "$(dirname "$0")/../../gpugi_b200"
make -s
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  case "$flags" in *--fmad*) ;; *) flags="--fmad=false $flags";; esac   # the product build's setting unless the variant chooses
  mkdir -p variants/$tag
  cp libgpugi_b200_host.so variants/$tag/
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr $flags -shared -o variants/$tag/libgpugi_b200.so csrc/gi_device.cu 2> variants/$tag/ptxas.log &
done
wait
for f in variants/*/ptxas.log; do echo "$f: $(grep -A2 'k_wf_traceILb0ELb0' $f | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"; done
