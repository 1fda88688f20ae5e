mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.log; tail -3 gpurun_out/bench_r1_a.log; cat gpurun_out/bench_r1_a.json
