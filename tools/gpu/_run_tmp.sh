CONFIGS="C2 C3 C4" ITERS=32 REPS=3 bash tools/gpu/ab_configs.sh cext9 cext10
export GPUGI_B200_LIBDIR=$(pwd)/gpugi_b200/variants/cext9
for T in 20 24 28; do for M in 12 16 20; do echo "T=$T M=$M"; GPUGI_WF_REFILL=$T GPUGI_WF_MIN_INNER=$M python tools/gpu/prof_batch.py --config C2 --iters 32 --repeat 3 --kernel-timing 2>/dev/null | grep '^{' | python -c "
import sys,json
r=min((json.loads(l) for l in sys.stdin), key=lambda r:r['ms_per_iteration']); print('   %.3f trace %.3f'%(r['ms_per_iteration'], r['kernel_ms_per_iteration']['trace']))"; done; done
