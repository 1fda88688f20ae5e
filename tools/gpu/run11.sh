#!/bin/bash
# trace-kernel experiment: parity tests, then the scheduler comparison (batch = the bench schedule)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/gpu/renderer_sweep.py --iters 24 --renderers ${1:-PT} --out gpurun_out/renderer_sweep_exp.json 2>&1 | grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('%-4s %-10s %-10s %7.3f ms/it  %7.1f Mpaths/s  %7.1f Mrays/s' % (r['renderer'], r['scheduler'], r['mode'], r['ms_per_iteration'], r['mpaths_per_s'], r['mrays_per_s']))
"
