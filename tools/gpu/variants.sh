#!/bin/bash
# Builds A/B variants of libgpugi_b200.so (CPU box; nvcc cross-compiles): tools/gpu/variants.sh tag "-DFLAG=0 ..." [tag flags ...]
# -> gpugi_b200/variants/libgpugi_b200_<tag>.so (git-ignored, travels with gpurun).  tools/gpu/run_variants.sh times them.
set -e
cd "$(dirname "$0")/../../gpugi_b200"
mkdir -p variants
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr $flags -shared -o variants/libgpugi_b200_$tag.so csrc/gi_device.cu 2> variants/ptxas_$tag.log &
done
wait
for f in variants/ptxas_*.log; do echo "$f: $(grep -A2 'k_wf_traceILb0ELb0' $f | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"; done
