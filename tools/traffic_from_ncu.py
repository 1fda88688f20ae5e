#!/usr/bin/env python
"""profiles/traffic.json from ncu --set full captures: per BASELINE config, what the steady-state launch of each kernel moved and what bound it.
    python tools/traffic_from_ncu.py C2 "source note" trace.ncu-rep [shade.ncu-rep ...]
bench.py reads profiles/traffic.json[<config>] for roofline.traffic / dram / l2 / issue (the live numbers come from CUDA events)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import load  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(x):
    return float(str(x).replace(",", ""))


def scale(v, unit, want):
    f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    return num(v) * f[unit] / (1.0 if want == "byte" else 1.0)


def main():
    cfg, note, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        data = json.load(open(path))
    except Exception:
        data = {}
    if "k_wf_trace_dram_bytes_per_launch" in data:  # round-1 flat file
        data = {}
    entry = {"source": note}
    for rep in reps:
        kernels, units = load(rep)
        for k in kernels:
            name = k["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0].strip()
            ms = scale(k["gpu__time_duration.sum"], units["gpu__time_duration.sum"], "ms")
            dram = scale(k["dram__bytes_read.sum"], units["dram__bytes_read.sum"], "byte") + scale(k["dram__bytes_write.sum"], units["dram__bytes_write.sum"], "byte")
            l2 = num(k["lts__t_sectors.sum"]) * 32.0
            entry[name + "_dram_bytes_per_launch"] = int(dram)
            entry[name + "_duration_ms_in_capture"] = ms
            entry[name + "_l2_bytes_per_launch"] = int(l2)
            entry[name + "_l2_gbs_in_capture"] = l2 / (ms * 1e-3) / 1e9
            entry[name + "_issue"] = {
                "issue_active_pct": num(k["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                "alu_pipe_pct": num(k["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]),
                "fma_pipe_pct": num(k["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]),
                "lanes_per_instruction": num(k["smsp__thread_inst_executed_per_inst_executed.ratio"]),
                "warps_active_pct": num(k["sm__warps_active.avg.pct_of_peak_sustained_active"]),
                "registers": int(num(k["launch__registers_per_thread"])),
                "rep": os.path.basename(rep)}
    data[cfg] = entry
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
