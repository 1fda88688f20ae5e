mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_r1_d.json 2> gpurun_out/bench_r1_d.log; tail -3 gpurun_out/bench_r1_d.log; cat gpurun_out/bench_r1_d.json
