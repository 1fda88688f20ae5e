#!/usr/bin/env python
"""Turns one round's raw GPU outputs (gpurun_out/) into the tracked evidence under profiles/:
    python tools/profile_pack.py <tag> "<title>"      e.g.  python tools/profile_pack.py l "PLOC topology + 4-wide nodes"
reads  gpurun_out/bench_r1_<tag>.json, bench_r1_<tag>_reftopo.json, launches_r1_<tag>.csv, ncu_r1<tag>_{trace,shade}.ncu-rep
writes profiles/r01_bench_wavefront_<tag>*.json, r01_launches_wavefront_<tag>.{csv,md}, r01_ncu_wavefront_<tag>.md, traffic.json"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag, title = sys.argv[1], sys.argv[2]
    go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    shutil.copy(os.path.join(go, "bench_r1_%s.json" % tag), os.path.join(pr, "r01_bench_wavefront_%s.json" % tag))
    if os.path.exists(os.path.join(go, "bench_r1_%s_reftopo.json" % tag)):
        shutil.copy(os.path.join(go, "bench_r1_%s_reftopo.json" % tag), os.path.join(pr, "r01_bench_wavefront_%s_reference_topology.json" % tag))
    shutil.copy(os.path.join(go, "launches_r1_%s.csv" % tag), os.path.join(pr, "r01_launches_wavefront_%s.csv" % tag))
    md = os.path.join(pr, "r01_ncu_wavefront_%s.md" % tag)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"),
                           "Round 1 (%s): ncu --set full of the steady-state k_wf_trace / k_wf_shade_pt waves -- %s -- bench.py --steps 24 --warmup 3" % (tag, title), md,
                           os.path.join(go, "ncu_r1%s_trace.ncu-rep" % tag), os.path.join(go, "ncu_r1%s_shade.ncu-rep" % tag)])
    # launch list shares
    with open(os.path.join(pr, "r01_launches_wavefront_%s.csv" % tag)) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        if len(name) > 60:
            name = name[:57] + "..."
        v, u = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        tot[name] += ms; cnt[name] += 1
    s = sum(tot.values())
    b = json.load(open(os.path.join(pr, "r01_bench_wavefront_%s.json" % tag)))
    km = b["roofline"]["kernel_ms_per_step"]; t = sum(km.values())
    out = ["# ncu launch list of `python bench.py --steps 24 --warmup 3 --no-cpu-baseline` (round 1, %s: %s)" % (tag, title), "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -c 600`; per-launch times are serialised and cold-cache, so only the SHARES are compared "
           "with the live CUDA-event timing of bench.py (`roofline.share_of_step`, `kernel_ms_per_step`).", "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
        out.append("| `%s` | %d | %.3f | %.1f %% |" % (k, cnt[k], v, 100 * v / s))
    out += ["", "Live timing in the bench (CUDA events per launch on the launching stream, %d steps): " % b["steps"] +
            ", ".join("%s %.3f ms (%.1f %%)" % (k, v, 100 * v / t) for k, v in km.items()) + "."]
    open(os.path.join(pr, "r01_launches_wavefront_%s.md" % tag), "w").write("\n".join(out) + "\n")
    # DRAM traffic per launch from the ncu summary
    vals = {}
    kernel = None
    for l in open(md):
        if l.startswith("## "):
            kernel = "trace" if "k_wf_trace" in l else "shade"
        if l.startswith("| gpu__time_duration.sum |"):
            p = [x.strip() for x in l.strip().strip("|").split("|")]
            vals[kernel + "_ms"] = float(p[1]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(p[2], 1.0)
        if l.startswith("| lts__t_sectors.sum |"):
            p = [x.strip() for x in l.strip().strip("|").split("|")]
            vals[kernel + "_l2"] = float(p[1]) * 32.0  # 32-byte sectors through the L2 tag stage
        if l.startswith("| dram__bytes_"):
            p = [x.strip() for x in l.strip().strip("|").split("|")]
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[p[2]]
            vals[kernel] = vals.get(kernel, 0.0) + float(p[1]) * scale
    json.dump({"k_wf_trace_dram_bytes_per_launch": int(vals["trace"]), "k_wf_shade_pt_dram_bytes_per_launch": int(vals["shade"]),
               "k_wf_trace_l2_bytes_per_launch": int(vals.get("trace_l2", 0)), "k_wf_shade_pt_l2_bytes_per_launch": int(vals.get("shade_l2", 0)),
               "k_wf_trace_ncu_launch_ms": vals.get("trace_ms"), "k_wf_shade_pt_ncu_launch_ms": vals.get("shade_ms"),
               "k_wf_trace_l2_gbs_in_capture": (vals.get("trace_l2", 0) / (vals["trace_ms"] * 1e-3) / 1e9) if vals.get("trace_ms") else None,
               "source": "profiles/r01_ncu_wavefront_%s.md (ncu --set full, steady-state wave of a 24-iteration batch: dram__bytes_read.sum + dram__bytes_write.sum; L2: lts__t_sectors.sum x 32 B)" % tag},
              open(os.path.join(pr, "traffic.json"), "w"), indent=1)
    print("\n".join(out[4:]))
    print({k: b[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches")}, b["roofline"]["achieved"], b["roofline"]["frac"], b["roofline"]["kernel_ms"])


if __name__ == "__main__":
    main()
