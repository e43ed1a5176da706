import sys, numpy as np
a = np.load(sys.argv[1]); b = np.load(sys.argv[2])
for l in range(8):
    A = set(zip(a["x%d"%l].tolist(), a["y%d"%l].tolist(), a["s%d"%l].tolist())); B = set(zip(b["x%d"%l].tolist(), b["y%d"%l].tolist(), b["s%d"%l].tolist()))
    print("level", l, len(A), len(B), "onlyA", sorted(A-B)[:12], "onlyB", sorted(B-A)[:12])
