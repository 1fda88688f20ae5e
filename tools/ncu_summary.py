#!/usr/bin/env python
"""Summarise one or more .ncu-rep captures (ncu --set full) as a markdown table for profiles/.
usage: python tools/ncu_summary.py title out.md rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALL = "smsp__average_warps_issue_stalled_"


def load(path):
    # a .csv is the saved output of `ncu -i capture.ncu-rep --page raw --csv` (made on the GPU box so that the large capture need not travel)
    out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    title, out_path, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    lines = ["# " + title, ""]
    for rep in reps:
        kernels, units = load(rep)
        for k in kernels:
            lines += ["## %s  (%s)" % (k.get("Kernel Name", "?"), rep.split("/")[-1]), "", "| metric | value | unit |", "|---|---|---|"]
            for w in WANT:
                if w in k:
                    lines.append("| %s | %s | %s |" % (w, k[w], units.get(w, "")))
            stalls = sorted(((float(v.replace(",", "")), n) for n, v in k.items() if n.startswith(STALL) and n.endswith("_per_issue_active.ratio") and v), reverse=True)
            if stalls:
                lines += ["", "Warp stall reasons (warps stalled per issue-active cycle, top 8):", "", "| reason | ratio |", "|---|---|"]
                for v, n in stalls[:8]:
                    lines.append("| %s | %.3f |" % (n[len(STALL):-len("_per_issue_active.ratio")], v))
            lines.append("")
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
