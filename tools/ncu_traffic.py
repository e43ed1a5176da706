#!/usr/bin/env python3
"""DRAM traffic per frame of each ORB stage from one `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum).
usage: ncu_traffic.py rep frames_per_launch out.json [first_n_launches]"""
import csv, json, subprocess, sys
rep, frames, outp = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
kn = hdr.index("Kernel Name"); ir = hdr.index("dram__bytes_read.sum"); iw = hdr.index("dram__bytes_write.sum"); it = hdr.index("gpu__time_duration.sum")
ii = hdr.index("smsp__inst_executed.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
stage_of = {"pyr_resize_kernel": "pyramid", "pyr_tma_kernel": "pyramid", "fast_cells_kernel": "fast", "octree_kernel": "octree", "orb_index_kernel": "index", "blur_kernel": "blur", "blur_tma_kernel": "blur",
            "orient_desc_kernel": "orient_desc"}
acc = {}
nmax = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 30
for r in rows[2:2 + nmax]:
    name = r[kn].split("(")[0].split("::")[-1].split("<")[0].strip().split(" ")[-1]
    st = stage_of.get(name)
    if not st: continue
    b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    e = acc.setdefault(st, {"dram_bytes_per_frame": 0.0, "ncu_us_per_frame": 0.0, "warp_inst_per_frame": 0.0, "launches": 0})
    e["dram_bytes_per_frame"] += b / frames; e["ncu_us_per_frame"] += float(r[it]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[it], 1.0) / frames; e["launches"] += 1
    e["warp_inst_per_frame"] += float(r[ii]) / frames
json.dump({"source": rep.split("/")[-1], "frames_per_launch": frames, "note": "ncu --set full --clock-control none, cold-cache serialised replays; one launch set",
           "stages": acc}, open(outp, "w"), indent=1)
print(json.dumps(acc, indent=1))
