#!/bin/bash
# GPU box, one GPU: the round's evidence for profiles/ -- bench lines of C2..C5, the ncu launch list of the bench command, and
# ncu --set full captures of the steady-state wavefront kernels of C2 (PT), C3 (BPT) and C4 (LPT).
#   tools/gpu/evidence_1gpu.sh <tag>     -> gpurun_out/<tag>_*
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${TAG}_bench_c2_n1.json 2> $O/${TAG}_bench_c2_n1.err
for c in C3 C4 C5; do python bench.py --config $c --steps 6 > $O/${TAG}_bench_$(echo $c | tr A-Z a-z)_n1.json 2> $O/${TAG}_bench_$c.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches_c2.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
# gpurun brings back at most 64 MiB: every capture is reduced to its raw page (CSV) on the box; only the C2 trace capture travels whole
trap 'for r in $O/${TAG}_ncu_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.csv 2>/dev/null; case $r in *c2_trace*) ;; *) rm -f $r;; esac; done; ls -la $O/${TAG}_*' EXIT
$NCU -k regex:k_wf_trace -s 20 -o $O/${TAG}_ncu_c2_trace python tools/gpu/prof_batch.py --config C2 --iters 26 > $O/${TAG}_ncu_c2_trace.log 2>&1
$NCU -k regex:k_wf_shade_pt -s 20 -o $O/${TAG}_ncu_c2_shade python tools/gpu/prof_batch.py --config C2 --iters 26 > $O/${TAG}_ncu_c2_shade.log 2>&1
$NCU -k regex:k_wf_shade_eye -s 24 -o $O/${TAG}_ncu_c3_shade_eye python tools/gpu/prof_batch.py --config C3 --iters 44 > $O/${TAG}_ncu_c3_shade_eye.log 2>&1
$NCU -k regex:k_wf_shade_light -s 24 -o $O/${TAG}_ncu_c3_shade_light python tools/gpu/prof_batch.py --config C3 --iters 44 > $O/${TAG}_ncu_c3_shade_light.log 2>&1
$NCU -k regex:k_wf_generate_light -s 24 -o $O/${TAG}_ncu_c3_generate_light python tools/gpu/prof_batch.py --config C3 --iters 44 > $O/${TAG}_ncu_c3_generate_light.log 2>&1
$NCU -k regex:k_wf_trace -s 70 -o $O/${TAG}_ncu_c3_trace python tools/gpu/prof_batch.py --config C3 --iters 44 > $O/${TAG}_ncu_c3_trace.log 2>&1
$NCU -k regex:k_wf_shade_light -s 20 -o $O/${TAG}_ncu_c4_shade_light python tools/gpu/prof_batch.py --config C4 --iters 26 > $O/${TAG}_ncu_c4_shade_light.log 2>&1
$NCU -k regex:k_wf_trace -s 20 -o $O/${TAG}_ncu_c4_trace python tools/gpu/prof_batch.py --config C4 --iters 26 > $O/${TAG}_ncu_c4_trace.log 2>&1
$NCU -k regex:k_wf_generate_light -s 20 -o $O/${TAG}_ncu_c4_generate_light python tools/gpu/prof_batch.py --config C4 --iters 26 > $O/${TAG}_ncu_c4_generate_light.log 2>&1
$NCU -k regex:k_wf_trace -s 20 -o $O/${TAG}_ncu_c5_trace python tools/gpu/prof_batch.py --config C5 --iters 26 > $O/${TAG}_ncu_c5_trace.log 2>&1
