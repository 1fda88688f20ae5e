python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2e_bench_c2_n1.json 2> gpurun_out/r2e_bench_c2.err
for c in C3 C4 C5; do python bench.py --config $c --steps 6 > gpurun_out/r2e_bench_$(echo $c | tr A-Z a-z)_n1.json 2> gpurun_out/r2e_bench_$c.err; done
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2e_bench_reference_c2.json 2>/dev/null
tail -c 700 gpurun_out/r2e_bench_c2_n1.json; echo; cat gpurun_out/r2e_bench_reference_c2.json | cut -c1-400
