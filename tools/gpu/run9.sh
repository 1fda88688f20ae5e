#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.log; tail -3 gpurun_out/bench_r1_f.log; cat gpurun_out/bench_r1_f.json
