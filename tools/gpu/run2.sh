mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.log; tail -3 gpurun_out/bench_r1_b.log; cat gpurun_out/bench_r1_b.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -30 gpurun_out/launches_r1.csv
ncu --set full --clock-control none --import-source on -k regex:k_pathtrace -s 4 -c 2 -o gpurun_out/prof_pt_r1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
