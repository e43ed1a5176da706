"""Runs SearchForInitialization (device-resident, 5000 x 5000 keypoints) a few times; run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from eorb_slam_b200 import api, synth
k1, d1, k2, d2, b = synth.make_keypoint_frame_pair(5000, 5000, 31)
prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
gm = api.GuidedMatcher(0, 0.9, True)
tk1 = torch.from_numpy(k1.view(np.uint8).reshape(-1).copy()).cuda(); tk2 = torch.from_numpy(k2.view(np.uint8).reshape(-1).copy()).cuda()
td1 = torch.from_numpy(d1).cuda(); td2 = torch.from_numpy(d2).cuda()
tprev0 = torch.from_numpy(prev).cuda(); tprev = tprev0.clone(); tm12 = torch.zeros(len(k1), dtype=torch.int32, device="cuda")
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    tprev.copy_(tprev0)
    n = gm.SearchForInitialization_device(tk1.data_ptr(), td1.data_ptr(), len(k1), tk2.data_ptr(), td2.data_ptr(), len(k2), b, tprev.data_ptr(), tm12.data_ptr(), 100)
print("nmatches", n)
