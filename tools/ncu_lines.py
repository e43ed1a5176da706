#!/usr/bin/env python3
"""Per-CUDA-source-line executed warp instructions of one kernel (needs -lineinfo and --import-source on).
usage: ncu_lines.py rep kernel_name [min_pct]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; fname = ""; items = []; nfun = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] != "":
        try:
            ie = int(r[hdr.index("Instructions Executed")]); sm = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        items.append((fname, int(r[0]), r[1].strip()[:100], ie, sm))
tot = sum(i[3] for i in items); st = sum(i[4] for i in items)
print("total warp instructions (all launches in report matching)", tot, "samples", st)
for f, ln, src, ie, sm in items:
    if ie >= tot * minpct / 100.0:
        print("%-16s %4d %11d %5.1f%%  smp %5.1f%% | %s" % (f, ln, ie, 100.0 * ie / tot, 100.0 * sm / max(st, 1), src))
