mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.log; tail -3 gpurun_out/bench_r1_c.log; cat gpurun_out/bench_r1_c.json
GI_WF_PROFILE=1 python tools/gpu/wf_sweep.py --iters 1 --only-wavefront 2>&1 | tail -45
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1_wf.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_wf.log 2>&1
tail -5 gpurun_out/launches_r1_wf.csv
bash tools/gpu/ncu_wf.sh r1c > /dev/null 2>&1
ls -la gpurun_out
