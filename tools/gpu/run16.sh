#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
python bench.py --gpus ${1:-4} --steps 20 --warmup 3 > gpurun_out/bench_r1_n${1:-4}.json 2> gpurun_out/bench_r1_n${1:-4}.log; tail -2 gpurun_out/bench_r1_n${1:-4}.log | cut -c1-200; cut -c1-400 gpurun_out/bench_r1_n${1:-4}.json; wc -l gpurun_out/bench_r1_n${1:-4}.json
python bench.py --impl reference --gpus ${1:-4} --steps 2 --warmup 3 2>/dev/null | cut -c1-300
