#!/usr/bin/env python3
"""Per-opcode executed-instruction histogram of one kernel from an .ncu-rep source page.
usage: ncu_opcodes.py rep kernel_name [launch-id]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern]
if len(sys.argv) > 3:
    cmd += ["--launch-skip", sys.argv[3], "--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several launches may be concatenated; take the first block
hdr = None; tot = collections.Counter(); thr = collections.Counter(); n = 0; samples = collections.Counter()
for r in rows:
    if r and r[0] == "Kernel Name":
        n += 1
        if n > 1: break
        continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    src = r[hdr.index("Source")].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    tot[op] += int(r[hdr.index("Instructions Executed")] or 0)
    thr[op] += int(r[hdr.index("Thread Instructions Executed")] or 0)
    samples[op] += int(r[hdr.index("# Samples")] or 0)
T = sum(tot.values()); S = sum(samples.values())
print("total warp instructions", T, "thread instructions", sum(thr.values()), "samples", S)
for op, c in tot.most_common(40):
    print("%-28s %12d %5.1f%%  lanes %.1f  samples %5.1f%%" % (op, c, 100.0 * c / T, thr[op] / max(c, 1), 100.0 * samples[op] / max(S, 1)))
