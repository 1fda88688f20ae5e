#!/bin/bash
# round-1 final check: GPU tests, bench line, ncu launch list of the bench command, --set full captures of the steady-state trace / shade waves
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_r1_i.json 2> gpurun_out/bench_r1_i.log; tail -2 gpurun_out/bench_r1_i.log; cut -c1-300 gpurun_out/bench_r1_i.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_i.csv python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r1_i.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 58 -c 1 -o gpurun_out/ncu_r1i_trace -f python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1i_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_shade -s 58 -c 1 -o gpurun_out/ncu_r1i_shade -f python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1i_shade.log 2>&1
ls -la gpurun_out/ | tail -8
