#!/bin/bash
# round-1 measurements on the re-clustered upload topology: smoke, bench line, ncu launch list of the bench command, --set full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_r1_n.json 2> gpurun_out/bench_r1_n.log; tail -2 gpurun_out/bench_r1_n.log; cut -c1-300 gpurun_out/bench_r1_n.json
python bench.py --steps 40 --warmup 5 --topology reference --no-cpu-baseline > gpurun_out/bench_r1_n_reftopo.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r1_n_reftopo.json
python bench.py --impl reference --steps 2 --warmup 3 2>/dev/null | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_n.csv python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r1_n.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 58 -c 1 -o gpurun_out/ncu_r1n_trace -f python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1n_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_shade -s 58 -c 1 -o gpurun_out/ncu_r1n_shade -f python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1n_shade.log 2>&1
ls -la gpurun_out/ | tail -6
