#!/bin/bash
# GPU box, one GPU, ~5 minutes: the GPU test suite with every parity statistic logged (GPUGI_B200_PARITY_LOG -> <tag>_parity_stats.jsonl) + smoke on
# the product build, then the bench lines of C2..C5 and, if gpugi_b200/variants/$AB_VARIANT exists (tools/gpu/variants.sh), an A/B against it on C4.
#   [AB_VARIANT=fmadon] tools/gpu/suite_and_bench.sh <tag>      -> gpurun_out/<tag>_*
TAG=${1:-r2g}
O=gpurun_out
mkdir -p $O
rm -f $O/${TAG}_parity_stats.jsonl
GPUGI_B200_PARITY_LOG=$O/${TAG}_parity_stats.jsonl timeout 200 python -m pytest tests -m gpu -x -q > $O/${TAG}_gpu_tests.log 2>&1
rc=$?
tail -3 $O/${TAG}_gpu_tests.log
timeout 60 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2 | tee $O/${TAG}_smoke.log
[ $rc -eq 0 ] || exit $rc
timeout 120 python bench.py > $O/${TAG}_bench_c2_n1.json 2> $O/${TAG}_bench_c2_n1.err
for c in C3 C4 C5; do timeout 150 python bench.py --config $c --steps 6 > $O/${TAG}_bench_$(echo $c | tr A-Z a-z)_n1.json 2> $O/${TAG}_bench_$c.err; done
AB=${AB_VARIANT:-fmadon}
if [ -d gpugi_b200/variants/$AB ]; then
  CONFIGS="C4" ITERS=24 REPS=2 tools/gpu/ab_configs.sh base $AB > $O/${TAG}_ab_${AB}_c4.txt 2>&1
  cat $O/${TAG}_ab_${AB}_c4.txt
fi
python - <<PY
import json
for c in ("c2", "c3", "c4", "c5"):
    try:
        d = json.loads(open("$O/${TAG}_bench_%s_n1.json" % c).read().strip().splitlines()[-1])
        print(c, round(d["value"], 1), round(d["ms_per_iteration"], 3), "e2e", round(d["e2e"]["value"], 1), "relmse", d.get("relmse", {}).get("value"), d["clocks"])
    except Exception as e:
        print(c, "failed", e)
PY
