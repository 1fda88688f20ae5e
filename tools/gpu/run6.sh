#!/bin/bash
# round-1 re-entry check: GPU tests, bench line, ncu launch list of the bench command, one --set full capture of k_wf_trace in the
# steady state of a batch (wave carrying 16 iterations) and of k_wf_shade_pt
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.log; tail -3 gpurun_out/bench_r1_e.log; cat gpurun_out/bench_r1_e.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_e.csv python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r1_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 58 -c 1 -o gpurun_out/ncu_r1e_trace -f python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1e_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_shade -s 58 -c 1 -o gpurun_out/ncu_r1e_shade -f python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1e_shade.log 2>&1
ls -la gpurun_out/
