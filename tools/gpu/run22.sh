#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/bench_r1_j_n$N.json 2> gpurun_out/bench_r1_j_n$N.log
tail -3 gpurun_out/bench_r1_j_n$N.log | cut -c1-300; cut -c1-600 gpurun_out/bench_r1_j_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus $N --steps 2 --warmup 3 2>/dev/null | cut -c1-300
