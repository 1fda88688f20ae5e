#!/bin/bash
mkdir -p gpurun_out
python bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_r1_n2.log; tail -5 gpurun_out/bench_r1_n2.log; cut -c1-700 gpurun_out/bench_r1_n2.json
