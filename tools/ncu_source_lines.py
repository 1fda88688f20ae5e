#!/usr/bin/env python
"""Per-source-line totals of an ncu capture made with --import-source on (kernels built with -lineinfo):
    python tools/ncu_source_lines.py capture.ncu-rep [top]
prints, per file:line, warp instructions executed, their share, the average active lanes and the stall samples."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file, hdr = None, None
    agg = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue
        d = dict(zip(hdr[4:], r[4:]))  # columns after the two Source columns
        try:
            inst = int(d["Instructions Executed"]); thr = int(d["Thread Instructions Executed"]); smp = int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        key = (cur_file, int(r[0]), r[1].strip()[:110])
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += inst; a[1] += thr; a[2] += smp
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    print("total warp instructions %d, thread instructions %d (%.2f lanes), samples %d" % (tot, sum(a[1] for a in agg.values()), sum(a[1] for a in agg.values()) / tot, tots))
    print("%-22s %6s %6s %6s  %s" % ("file:line", "inst%", "lanes", "smp%", "source"))
    for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s %6.2f %6.1f %6.2f  %s" % ("%s:%d" % (f, ln), 100.0 * a[0] / tot, a[1] / max(a[0], 1), 100.0 * a[2] / tots, src))


if __name__ == "__main__":
    main()
