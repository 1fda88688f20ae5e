#!/bin/bash
# GPU box with N GPUs (gpurun --gpus N): bench lines of the BASELINE configs at N ranks + the multi-GPU tests.
#   tools/gpu/evidence_ngpu.sh <tag> <N> "<configs>"     -> gpurun_out/<tag>_bench_<config>_n<N>.json
TAG=$1; N=$2; CFGS=${3:-"C2 C3"}
O=gpurun_out
mkdir -p $O
nvidia-smi -L | head -8
[ -n "$SKIP_TESTS" ] || python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3 | tee $O/${TAG}_multi_tests_n$N.log
for c in $CFGS; do
  lc=$(echo $c | tr A-Z a-z)
  if [ $c = C2 ]; then S=20; else S=6; fi
  python bench.py --gpus $N --config $c --steps $S > $O/${TAG}_bench_${lc}_n$N.json 2> $O/${TAG}_bench_${lc}_n$N.err
  tail -c 600 $O/${TAG}_bench_${lc}_n$N.json | head -c 300; echo
done
