#!/usr/bin/env python3
"""Resident ORB step with the launch sets spread over L extractor handles on L CUDA streams (concurrent launch sets fill each
other's issue slots / tails).  usage: lanes_probe.py [frames] -> frames/s per (lanes, chunk)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from eorb_slam_b200 import api, synth
import bench

nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W, H = 752, 480
frames = bench.make_batch(nfr, 0)
d_frames = torch.from_numpy(frames).cuda()
main = torch.cuda.current_stream()
for lanes, chunk in [(1, 1024), (2, 1024), (2, 512), (4, 512), (4, 256), (1, 512), (8, 256), (2, 2048) if nfr >= 4096 else (1, 256)]:
    p = api.ORBxParams()
    exs = [api.ORBextractor(p, 0, chunk) for _ in range(lanes)]
    cap = exs[0].cap
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    for e, s in zip(exs, streams):
        e.set_stream(s.cuda_stream)
    d_kps = torch.empty(nfr * cap * 28, dtype=torch.uint8, device="cuda"); d_desc = torch.empty(nfr * cap * 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(nfr, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(nfr, dtype=torch.int32, device="cuda")

    def step():
        for s in streams:
            s.wait_stream(main)
        for k, f0 in enumerate(range(0, nfr, chunk)):
            nb = min(chunk, nfr - f0)
            exs[k % lanes].extract_batch_raw(d_frames.data_ptr() + f0 * W * H, nb, W, H, W, W * H, (0, 1000), True,
                                             d_kps.data_ptr() + f0 * cap * 28, d_desc.data_ptr() + f0 * cap * 32, cap,
                                             d_n.data_ptr() + f0 * 4, d_mono.data_ptr() + f0 * 4, device=True)
        for s in streams:
            main.wait_stream(s)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(5):
        step()
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("lanes %d chunk %4d: %.3f ms/step  %.0f frames/s  (%.3f us/frame)  kps %d" % (lanes, chunk, ms, nfr / ms * 1e3, ms / nfr * 1e3, int(d_n.sum().item())), flush=True)
    del exs, d_kps, d_desc
    torch.cuda.empty_cache()
