#!/usr/bin/env python
"""profiles/r02_sass_k_wf_trace.txt: SASS of one kernel of the in-tree device library with a static opcode histogram.
   python tools/sass_listing.py [mangled kernel name] > profiles/r02_sass_k_wf_trace.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kernel = sys.argv[1] if len(sys.argv) > 1 else "_Z10k_wf_traceILb0ELb0EEvN3gik7WfFrameEPy"
out = subprocess.run(["cuobjdump", "-sass", "-fun", kernel, os.path.join(ROOT, "gpugi_b200", "libgpugi_b200.so")], capture_output=True, text=True).stdout
body, hist = [], collections.Counter()
for line in out.splitlines():
    if re.match(r"^\s+/\* 0x[0-9a-f]+ \*/\s*$", line):          # second encoding word
        continue
    m = re.match(r"^(\s+/\*[0-9a-f]{4,}\*/\s+.*?;)\s+/\* 0x[0-9a-f]+ \*/\s*$", line)
    if m:
        line = m.group(1)
        ins = line.split("*/", 1)[1].split()
        op = ins[1] if ins[0].startswith("@") else ins[0]
        hist[op.split(".")[0].rstrip(";")] += 1
    if "Function :" in line or ".headerflags" in line or m:
        body.append(line)
print("# SASS of k_wf_trace<false,false> (sm_100a), round-2 final build (--fmad=false): cuobjdump -sass gpugi_b200/libgpugi_b200.so, encodings stripped (tools/sass_listing.py)")
print("# static opcode histogram (whole kernel incl. the exact reference-order re-trace and the retire / refill code):")
for op, n in hist.most_common(42):
    print("#   %-10s %d" % (op, n))
print("# loads of the inner steps: three LDG.E.ENL2.256 (one 96-byte node record); no PRMT in the slab test; FMNMX3 = three-input min / max")
print("\n".join(body))
