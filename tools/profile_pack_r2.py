#!/usr/bin/env python
"""Turns the raw outputs of tools/gpu/evidence_1gpu.sh <tag> (gpurun_out/<tag>_*) into the tracked round-2 evidence under profiles/:
    python tools/profile_pack_r2.py <tag>
bench lines r02_bench_c{2..5}_n1.json, launch list r02_launches_c2.{csv,md}, ncu summaries r02_ncu_c{2..5}.md, traffic.json."""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def run(*a):
    subprocess.check_call([sys.executable] + list(a), stdout=subprocess.DEVNULL)


def main():
    tag = sys.argv[1]
    g = lambda name: os.path.join(GO, "%s_%s" % (tag, name))
    for c in ("c2", "c3", "c4", "c5"):
        shutil.copy(g("bench_%s_n1.json" % c), os.path.join(PR, "r02_bench_%s_n1.json" % c))
    shutil.copy(g("launches_c2.csv"), os.path.join(PR, "r02_launches_c2.csv"))
    summ, traf = os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(ROOT, "tools", "traffic_from_ncu.py")
    sets = {
        "c2": ("steady-state k_wf_trace / k_wf_shade_pt waves of C2 (prof_batch.py --config C2 --iters 26, launch 20 of each kernel)", ["c2_trace", "c2_shade"]),
        "c3": ("steady-state BPT waves of C3 (prof_batch.py --config C3 --iters 44: shade / generate kernels launch 24, k_wf_trace launch 70)",
               ["c3_trace", "c3_shade_eye", "c3_shade_light", "c3_generate_light"]),
        "c4": ("steady-state LPT waves of C4 (prof_batch.py --config C4 --iters 26, launch 20)", ["c4_trace", "c4_shade_light", "c4_generate_light"]),
        "c5": ("steady-state k_wf_trace wave of C5 (prof_batch.py --config C5 --iters 26, launch 20)", ["c5_trace"]),
    }
    for c, (what, reps) in sets.items():
        files = [g("ncu_%s.csv" % r) for r in reps]
        run(summ, "Round 2 (final kernels): ncu --set full of the " + what + "; tools/gpu/evidence_1gpu.sh", os.path.join(PR, "r02_ncu_%s.md" % c), *files)
        run(traf, c.upper(), "profiles/r02_ncu_%s.md (ncu --set full, one steady-state launch per kernel: dram__bytes_read.sum + dram__bytes_write.sum; L2: lts__t_sectors.sum x 32 B)" % c, *files)
    lines = [l for l in open(os.path.join(PR, "r02_launches_c2.csv")) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v, u = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
        tot[name] += v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        cnt[name] += 1
    T = sum(tot.values())
    out = ["# Round 2: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (C2), `--metrics gpu__time_duration.sum --clock-control none`", "",
           "Per-launch times under ncu are cold-cache and serialised; what must agree with the live timing (bench line `roofline.share_of_step`, CUDA events) is each kernel's SHARE.",
           "(`k_wf_trace<1, 1>` / `<0, 1>` are the two untimed counter passes of the roofline: reference-order walk and GI_OPT_COUNT_STEPS.)", "",
           "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out.append("| `%s` | %d | %.3f | %.1f %% |" % (k[:70], cnt[k], v, 100 * v / T))
    open(os.path.join(PR, "r02_launches_c2.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:12]))


if __name__ == "__main__":
    main()
