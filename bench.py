#!/usr/bin/env python3
"""bench.py — headline benchmark of the EORB-SLAM front-end hot path on B200.

Contract (one JSON line on stdout from rank 0):
  metric      orb_frames_per_sec: ORB extraction (752x480, nFeatures=1000, 8 levels, 1.2, FAST 20/7) over a batch
              of 4096 synthetic frames per GPU (BASELINE.json configs[2]); a step = one pass over the batch.
  value       frames/s with the frames resident in HBM (device-pointer C-ABI call), CUDA-event timed, max over ranks
  e2e         same metric through the host-buffer C-ABI call (pinned host frames in, keypoints+descriptors out;
              H2D and D2H inside the timed region)
  roofline    dominant kernel of the step vs the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own ORBextractor.cc compiled unmodified (oracle/_ref/libref.so, kind "reference"; the
              oracle port when that library is absent) on the box's host cores, bounded sample; the event and Hamming
              legs carry their own cpu_baseline (reference EventConversion.cc / threaded popcount)
  parity_checked  outside the timed regions: every frame's keypoint count and the first frames' full keypoint +
              descriptor bytes against the oracle; all 2000 match records of configs[3] against a threaded CPU best-2
  extra       event-frame Mev/s (configs[1], aggregate over ranks) and Hamming Gmatch/s (configs[3]) with their rooflines,
              strong scaling of configs[2] (4096 frames split over the ranks), PCIe link + solo/concurrent H2D per GPU
`--impl reference` times the reference's CPU implementation (libref.so, all host threads) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H = 752, 480
ORB_KW = dict(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, edge_th=19)
FRAMES_PER_GPU = 4096
CHUNK = 4096                     # frames per launch set of the device-resident path (the host path pipelines 128-frame slots)
UNIQUE_FRAMES = 256              # generated frames; the rest are circular shifts of these (all distinct)
# algorithmic bytes per frame (SURVEY.md §8d / BASELINE.md §2 / DESIGN.md)
LEVELS = [(752, 480), (627, 400), (522, 333), (435, 278), (363, 231), (302, 193), (252, 161), (210, 134)]
PIX = [w * h for w, h in LEVELS]
ALGO_BYTES = {
    "pyramid": sum(PIX[l - 1] + PIX[l] for l in range(1, 8)),      # 1 845 634
    "fast": sum(PIX),                                             # 1 117 367 (+8 B/candidate, ignored)
    "blur": 2 * sum(PIX),                                         # 2 234 734
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples SM clock and throttle reasons with NVML while the timed region runs"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _traffic():
    """DRAM bytes per frame of each stage from the committed `ncu --set full` capture at the TIMED launch-set size (profiles/r02_traffic.json:
    1024 frames per launch, tools/ncu_traffic.py on gpurun_out/prof_r02c.ncu-rep)"""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(p):
        d = json.load(open(p))
        _traffic.frames_per_launch = d.get("frames_per_launch")
        _traffic.inst = {k: v.get("warp_inst_per_frame") for k, v in d["stages"].items()}
        return {k: v["dram_bytes_per_frame"] for k, v in d["stages"].items()}, d.get("source")
    _traffic.inst = {}
    return {}, None


def _bind_to_gpu_cpus(index):
    """one process per GPU: run on the CPUs NVML reports as local to this GPU so that the pinned batch and the
    library's pinned staging land on the GPU's NUMA node (H2D/D2H do not cross the socket interconnect)"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))[:4]
    except Exception as e:   # best effort: affinity is an optimisation, not a requirement
        return "unbound (%s)" % type(e).__name__


def make_batch(n, seed0):
    from eorb_slam_b200 import synth
    return synth.make_frames(n, seed0=seed0, w=W, h=H, unique=min(UNIQUE_FRAMES, n))


# ------------------------------------------------------------------------------------------------ reference arm
WORKLOAD = "configs[2]: ORB 752x480 nFeatures=1000 8 levels 1.2 FAST 20/7, %d frames per GPU per step"


def _config(nfr):
    """the workload description, identical in both arms"""
    return {"workload": WORKLOAD % nfr, "frames_per_gpu": nfr,
            "l2_policy": "inputs larger than L2 (%.0f MB of frames per step per GPU)" % (nfr * W * H / 1e6),
            "partition": "by frame, no collective"}


def _cpu_reference_extractor():
    """-> (batch function, kind, description): the reference's own code when oracle/_ref/libref.so is present"""
    try:
        import ref_lib as R
        if R.available():
            R.lib()
            return R.orb_extract_batch_mt, "reference", ("oracle/_ref/libref.so = /root/reference/src/ORBextractor.cc compiled unmodified "
                                                        "(OpenCV primitives: the cv2-pinned oracle primitives), glibc malloc")
    except Exception:
        pass
    import oracle_lib as O
    return O.orb_extract_batch_mt, "port", "CPU oracle port of src/ORBextractor.cc (libref.so absent)"


def run_reference(args):
    """The reference's CPU implementation of the path, all host threads, the SAME config as our arm (4096 frames per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    fn, kind, what = _cpu_reference_extractor()
    cores = os.cpu_count() or 1
    n = args.frames
    frames = make_batch(n, 0)
    for _ in range(args.warmup):
        fn(frames, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(frames, cores)
    dt = time.perf_counter() - t0
    fps = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "orb_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": _config(n),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": "%d frames per step (the whole batch), %d steps, std::thread pool over frames, one extractor per thread; %s"
                                   % (n, args.steps, what)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


def cv2_assisted_baseline(frames, cores):
    """BASELINE.md section 3, line 2: how fast could the reference's CPU path be with OpenCV's own SIMD kernels under it?
    Dense stages (resize, copyMakeBorder, FAST, GaussianBlur) through the cv2 wheel on whole levels, one worker PROCESS per core
    (no GIL), plus the non-dense remainder (grid bookkeeping, octree, orientation, BRIEF, assembly) taken from the oracle port:
    its whole-frame time minus its own dense primitives on the same levels.  A composed estimate, reported as such."""
    import multiprocessing as mp
    import oracle_lib as O
    n = min(len(frames), 16 * cores)
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cv2_dense_frame, [frames[i] for i in range(min(cores, n))])          # warm
        t0 = time.perf_counter()
        pool.map(_cv2_dense_frame, [frames[i] for i in range(n)], chunksize=max(1, n // (4 * cores)))
        dense_wall = time.perf_counter() - t0
    cv_dense_ms = dense_wall * cores / n * 1e3                   # per frame per core
    # the oracle's split on one core
    orc = O.OrbOracle(**{k: ORB_KW[k] for k in ("nfeatures", "scale_factor", "nlevels", "ini_th", "min_th", "edge_th")}, im_w=W, im_h=H)
    k = 4
    t0 = time.perf_counter()
    for i in range(k):
        orc.extract(frames[i])
    whole_ms = (time.perf_counter() - t0) / k * 1e3
    t0 = time.perf_counter()
    for i in range(k):
        prev = frames[i]
        for l in range(ORB_KW["nlevels"]):
            if l:
                prev = O.resize_linear(prev, LEVELS[l][0], LEVELS[l][1])
            O.border_reflect101(prev, 19); O.fast(prev, ORB_KW["ini_th"], True); O.gauss5(prev)
    port_dense_ms = (time.perf_counter() - t0) / k * 1e3
    rest_ms = max(whole_ms - port_dense_ms, 0.0)
    per_frame_ms = cv_dense_ms + rest_ms
    return {"value": cores * 1e3 / per_frame_ms, "unit": "frames/s", "cores": cores, "kind": "cv2-assisted (composed estimate)",
            "ms_per_frame_per_core": {"cv2_dense": cv_dense_ms, "oracle_rest": rest_ms, "oracle_whole": whole_ms, "oracle_dense": port_dense_ms},
            "sample": "%d frames through cv2 resize/copyMakeBorder/FAST/GaussianBlur on whole levels in %d worker processes; remainder = oracle "
                      "whole-frame time minus its dense primitives, one core" % (n, cores)}


def _cv2_dense_frame(img):
    import cv2
    cv2.setNumThreads(1)
    fast = cv2.FastFeatureDetector_create(threshold=ORB_KW["ini_th"], nonmaxSuppression=True)
    prev = img
    nk = 0
    for l in range(ORB_KW["nlevels"]):
        if l:
            prev = cv2.resize(prev, LEVELS[l], interpolation=cv2.INTER_LINEAR)
        cv2.copyMakeBorder(prev, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
        nk += len(fast.detect(prev, None))
        cv2.GaussianBlur(prev, (5, 5), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    return nk


# ------------------------------------------------------------------------------------------------ extras
def _events_cpu_baseline(ev, per, w, h, sigma, mode, Tcw, depth, K, nwin_sample):
    """the reference's EventConversion.cc (libref.so, unmodified) over consecutive windows, a std::thread pool over windows;
    falls back to the oracle port window by window on one core when libref is absent"""
    cores = os.cpu_count() or 1
    n = nwin_sample * per
    try:
        import ref_lib as R
        if R.available() and hasattr(R.lib(), "ref_ev_accumulate_batch_mt"):
            T = np.ascontiguousarray(Tcw, np.float32).reshape(16) if Tcw is not None else None
            Kc = np.ascontiguousarray(K, np.float32) if K is not None else np.zeros(4, np.float32)
            evs = np.ascontiguousarray(ev[:n])
            R.lib().ref_ev_accumulate_batch_mt(R._p(evs), min(n, cores * per), per, w, h, sigma, mode, R._p(T), depth, R._p(Kc), cores)
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 6.0:     # bounded sample: about 6 s of CPU work on all cores
                R.lib().ref_ev_accumulate_batch_mt(R._p(evs), n, per, w, h, sigma, mode, R._p(T), depth, R._p(Kc), cores)
                reps += 1
            dt = time.perf_counter() - t0
            return {"value": reps * n / dt / 1e6, "unit": "Mev/s", "cores": cores, "kind": "reference",
                    "sample": "%d x %d windows x %d events through EvImConverter::%s (normalised u8, libref.so = EventConversion.cc unmodified), "
                              "thread pool over windows, %.1f s" % (reps, nwin_sample, per, "ev2im_gauss" if mode == 1 else "ev2mci_gg_f", dt)}
    except Exception:
        pass
    import oracle_lib as O
    k = max(1, min(nwin_sample, 16))
    t0 = time.perf_counter()
    for i in range(k):
        O.ev_accumulate(ev[i * per:(i + 1) * per], w, h, sigma, mode=mode, Tcw=Tcw, depth=depth, K=K, normalize=True)
    dt = time.perf_counter() - t0
    return {"value": k * per / dt / 1e6, "unit": "Mev/s", "cores": 1, "kind": "port", "sample": "%d windows x %d events, oracle port, one core" % (k, per)}


def bench_events(api, torch, dev, steps, warmup, world=1, rank=0, dist=None, cpu=True):
    """configs[1]: DAVIS240 stream, fixed 2000-event windows -> Gaussian event frames (+ running normalisation).
    Windows are partitioned over the ranks (every rank owns 512 windows of its own stream, no collective); the
    reported value is the aggregate over ranks, timed as the max over ranks."""
    from eorb_slam_b200 import synth
    nwin, per, w, h = 512, 2000, 240, 180
    ev = synth.make_events(nwin * per, seed=1 + 1000 * rank, w=w, h=h)
    cv = api.EvImConverter(dev, nwin, nwin * per, w, h)
    st = torch.cuda.current_stream().cuda_stream
    cv.set_stream(st)
    d_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1)).cuda()
    d_img = torch.empty(nwin * h * w, dtype=torch.float32, device="cuda")
    d_u8 = torch.empty(nwin * h * w, dtype=torch.uint8, device="cuda")
    p = cv.make_params(api.EV_GAUSS, w, h, 1.0, False, api.NORM_RUNNING)
    offs = np.arange(nwin + 1, dtype=np.int64) * per
    for _ in range(warmup):
        cv.accumulate_batch_device(d_ev.data_ptr(), offs, p, d_img.data_ptr(), d_u8.data_ptr())
    torch.cuda.synchronize()
    l0 = cv.launch_count()
    t = api.CudaTimer()
    t.start(st)
    for _ in range(steps):
        cv.accumulate_batch_device(d_ev.data_ptr(), offs, p, d_img.data_ptr(), d_u8.data_ptr())
    t.stop(st)
    ms = t.elapsed_ms() / steps
    ms_rank = ms
    if world > 1:
        tm = torch.tensor([ms], device="cuda"); dist.all_reduce(tm, op=dist.ReduceOp.MAX); ms = float(tm.item())
    launches = cv.launch_count() - l0
    nev = nwin * per
    algo_bytes = nev * 24 + nwin * (4 * w * h + w * h)
    peak, _ = _peaks()
    out = {"metric": "event_frames_mev_per_s", "value": world * nev / ms / 1e3, "unit": "Mev/s", "ms_per_step": ms, "n_gpus": world,
           "scaling": "weak", "per_gpu_mev_per_s": nev / ms_rank / 1e3,
           "workload": "configs[1]: %d windows x %d events per GPU, 240x180, sigma 1 (49 taps/event), normalised to u8; windows partitioned "
                       "over the ranks, no collective" % (nwin, per),
           "atomic_adds_per_s": nev * 49 / (ms * 1e-3), "gpu_launches": launches,
           "roofline": {"bound": "hbm", "achieved": algo_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": algo_bytes / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                        "note": "7x7 splat accumulates in shared memory (int32 fixed point, native ATOMS.ADD), frame written once; "
                                "bounded by instruction issue + smem atomics, not HBM; see atomic_adds_per_s"}}
    cv.set_stream(None)
    # end to end through the host entry point: pinned host events in (24 B each), u8 event frames back in pinned host memory, copies inside
    # the timed region (eorb_ev_accumulate_batch: one H2D, the batch kernels, one D2H)
    try:
        h_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1).copy()).pin_memory()
        h_u8 = torch.empty(nwin * h * w, dtype=torch.uint8).pin_memory()
        for _ in range(3):
            cv.accumulate_batch(h_ev.data_ptr(), offs, p, None, h_u8.data_ptr())
        if world > 1:
            dist.barrier()
        nrep = max(steps, 10)
        t0 = time.perf_counter()
        for _ in range(nrep):
            cv.accumulate_batch(h_ev.data_ptr(), offs, p, None, h_u8.data_ptr())
        ems = (time.perf_counter() - t0) * 1e3 / nrep
        if world > 1:
            tm = torch.tensor([ems], device="cuda"); dist.all_reduce(tm, op=dist.ReduceOp.MAX); ems = float(tm.item())
        same = bool(torch.equal(h_u8, d_u8.cpu()))
        out["e2e"] = {"value": world * nev / ems / 1e3, "unit": "Mev/s", "ms_per_step": ems, "h2d_bytes_per_step": nev * 24, "d2h_bytes_per_step": nwin * h * w,
                      "call": "eorb_ev_accumulate_batch (pinned host events in, u8 frames out)", "frames_equal_resident_path": same}
    except Exception as e:
        out["e2e"] = {"error": repr(e)}
    if cpu and rank == 0 and world == 1:
        out["cpu_baseline"] = _events_cpu_baseline(ev, per, w, h, 1.0, 1, None, 1.0, None, nwin)
    # configs[1] as ONE device pipeline: the windows above -> u8 event frames -> single-level event extractor (N = 400, FAST 0/0,
    # keypoints only, EvETHZ.yaml:185-199) on the frames where they lie in HBM
    try:
        exe = api.ORBextractor(api.ORBxParams(400, 1.0, 1, 0, 0, 9, (w, h)), dev, nwin)
        exe.set_stream(st)
        ecap = exe.cap
        d_k = torch.empty(nwin * ecap * 28, dtype=torch.uint8, device="cuda"); d_d = torch.empty(nwin * ecap * 32, dtype=torch.uint8, device="cuda")
        d_nn = torch.zeros(nwin, dtype=torch.int32, device="cuda"); d_mm = torch.zeros(nwin, dtype=torch.int32, device="cuda")
        cv.set_stream(st)

        def chain():
            cv.accumulate_batch_device(d_ev.data_ptr(), offs, p, d_img.data_ptr(), d_u8.data_ptr())
            exe.extract_batch_raw(d_u8.data_ptr(), nwin, w, h, w, w * h, (0, 1000), False, d_k.data_ptr(), d_d.data_ptr(), ecap, d_nn.data_ptr(),
                                  d_mm.data_ptr(), device=True)
        for _ in range(3):
            chain()
        torch.cuda.synchronize()
        t.start(st)
        for _ in range(10):
            chain()
        t.stop(st)
        msc = t.elapsed_ms() / 10
        out["chain_events_to_keypoints"] = {"value": nwin / (msc * 1e-3), "unit": "windows/s", "ms_per_step": msc, "mev_per_s": nev / msc / 1e3,
                                            "keypoints_per_window": float(d_nn.float().mean().item()),
                                            "workload": "configs[1]: %d windows x %d events -> 240x180 u8 event frames -> single-level ORB "
                                                        "(N=400, FAST 0/0, keypoints only), all in HBM" % (nwin, per)}
        cv.set_stream(None); exe.set_stream(None)
    except Exception as e:
        out["chain_events_to_keypoints"] = {"error": repr(e)}
    # configs[4]: MVSEC-shaped 346x260 frames, 50k events per window, motion compensated with a per-window rotation
    # (ev2mci_gg_f, SE3 warp in double per event), cv::normalize(MINMAX) to u8.  The 360 KB frame does not fit one SM's
    # shared memory, so it is split into row bands (every band scans the window's events).
    try:
        nw2, per2, w2, h2 = 74, 50000, 346, 260   # 74 windows x 2 row bands = one block per SM
        ev2 = synth.make_events(nw2 * per2, seed=2, w=w2, h=h2)
        cv2_ = api.EvImConverter(dev, nw2, nw2 * per2, w2, h2)
        cv2_.set_stream(st)
        d_ev2 = torch.from_numpy(ev2.view(np.uint8).reshape(-1)).cuda()
        d_img2 = torch.empty(nw2 * h2 * w2, dtype=torch.float32, device="cuda")
        d_u82 = torch.empty(nw2 * h2 * w2, dtype=torch.uint8, device="cuda")
        dt = float(ev2["ts"][per2 - 1] - ev2["ts"][0])
        T = synth.rotation_tcw(np.array([0.5, -0.7, 1.5]) * dt)
        p2 = cv2_.make_params(api.EV_SE3, w2, h2, 1.0, False, api.NORM_MINMAX, Tcw=T, medDepth=1.0, camera=(226.38, 226.15, 173.65, 133.73))
        offs2 = np.arange(nw2 + 1, dtype=np.int64) * per2
        for _ in range(3):
            cv2_.accumulate_batch_device(d_ev2.data_ptr(), offs2, p2, d_img2.data_ptr(), d_u82.data_ptr())
        torch.cuda.synchronize()
        t.start(st)
        for _ in range(10):
            cv2_.accumulate_batch_device(d_ev2.data_ptr(), offs2, p2, d_img2.data_ptr(), d_u82.data_ptr())
        t.stop(st)
        ms2 = t.elapsed_ms() / 10
        out["mc_346x260"] = {"value": nw2 * per2 / ms2 / 1e3, "unit": "Mev/s", "ms_per_step": ms2, "n_gpus": world,
                             "workload": "configs[4]: %d windows x %d events per GPU, 346x260, SE3 motion compensation, sigma 1, MINMAX u8" % (nw2, per2)}
        if cpu and rank == 0 and world == 1:
            out["mc_346x260"]["cpu_baseline"] = _events_cpu_baseline(ev2, per2, w2, h2, 1.0, 2, T, 1.0, (226.38, 226.15, 173.65, 133.73), nw2)
        cv2_.set_stream(None)
    except Exception as e:   # the second workload never invalidates the first
        out["mc_346x260"] = {"error": repr(e)}
    if world > 1:   # aggregate over ranks, max-over-ranks time (outside the try: every rank reaches this collective)
        mine = out["mc_346x260"].get("ms_per_step", 0.0)
        tm = torch.tensor([mine], device="cuda"); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        if "ms_per_step" in out["mc_346x260"]:
            out["mc_346x260"]["ms_per_step"] = float(tm.item())
            out["mc_346x260"]["value"] = world * out["mc_346x260"]["value"] * mine / float(tm.item())
    return out


def bench_hamming(api, torch, dev, steps, warmup, world, rank, dist, cpu=True):
    """configs[3]: 2000 queries vs a 16M-row database, row-sharded over the ranks, all-gather of per-shard best-2.
    Outside the timed region rank 0 rebuilds every shard (same device generator seeds), runs the threaded CPU best-2 over the
    whole database and compares all 2000 (distance, index, second distance, accepted) records with the merged GPU result:
    that run is both the parity check of every cross-rank merge and the Hamming leg's cpu_baseline."""
    nq, ndb_total = 2000, 16 * 1024 * 1024
    per = ndb_total // world
    g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
    d_db = torch.randint(0, 256, (per, 32), dtype=torch.uint8, device="cuda", generator=g)
    gq = torch.Generator(device="cuda"); gq.manual_seed(99)
    d_q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda", generator=gq)
    if rank == 0:
        d_q[:1000] = d_db[:1000]            # planted exact matches in shard 0
    if world > 1:
        dist.broadcast(d_q, 0)
    m = api.ORBmatcher(0.7, True, dev)
    st = torch.cuda.current_stream().cuda_stream
    m.set_stream(st)
    m.set_db_device(d_db.data_ptr(), per, rank * per)
    out = torch.empty(nq * 16, dtype=torch.uint8, device="cuda")
    # communicator for the C-ABI sharded call: rank 0 creates the ncclUniqueId, torch.distributed only carries its 128 bytes
    uid = torch.from_numpy(api.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
    if world > 1:
        dist.broadcast(uid, 0)
    comm = api.nccl_comm_init_rank(world, uid.cpu().numpy(), rank, dev)

    def step():
        # shard scan + ncclAllGather (nq x 16 B per rank) + merge, one C-ABI call on the matcher's stream
        m.search_sharded(d_q.data_ptr(), nq, comm, world, out.data_ptr(), th=50)

    def timed(engine):
        m.set_engine(engine)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0_ = m.launch_count()
        t = api.CudaTimer()
        t.start(st)
        for _ in range(steps):
            step()
        t.stop(st)
        ms_ = t.elapsed_ms() / steps
        if world > 1:
            tm = torch.tensor([ms_], device="cuda"); dist.all_reduce(tm, op=dist.ReduceOp.MAX); ms_ = float(tm.item())
        return ms_, m.launch_count() - l0_, out.cpu().numpy().tobytes()

    # the POPC engine (the reference's DescriptorDistance as is) first, then the default: AUTO = the tensor-core engine at this size
    ms_popc, _, bytes_popc = timed(m.HAMMING_POPC)
    ms, nlaunch, bytes_main = timed(m.HAMMING_AUTO)
    engine_used = m.last_engine()
    l0 = m.launch_count() - nlaunch
    res = np.frombuffer(out.cpu().numpy().tobytes(), dtype=np.dtype([("d", "<i4"), ("i", "<i4"), ("s", "<i4"), ("a", "<i4")]))
    ok = bool((res["d"][:1000] == 0).all() and (res["i"][:1000] == np.arange(1000)).all())
    pairs = nq * ndb_total
    popc_peak = api.probe_popc_rate(dev)
    # end to end through the host entry point (one GPU: the index is resident, as a database is; the 2000 queries come from host memory and the
    # 2000 eorb_match records go back to it inside the timed region)
    e2e_h = None
    if world == 1:
        try:
            q_h = d_q.cpu().numpy()
            for _ in range(2):
                rec = m.search(q_h)
            nrep = max(steps, 5)
            t0 = time.perf_counter()
            for _ in range(nrep):
                rec = m.search(q_h)
            hms = (time.perf_counter() - t0) * 1e3 / nrep
            same_h = bool(np.array_equal(rec["best_dist"], res["d"]) and np.array_equal(rec["best_idx"], res["i"]) and np.array_equal(rec["second_dist"], res["s"])
                          and np.array_equal(rec["accepted"].astype(np.int32), res["a"]))
            e2e_h = {"value": pairs / (hms * 1e-3) / 1e9, "unit": "Gmatch/s", "ms_per_step": hms, "h2d_bytes_per_step": nq * 32, "d2h_bytes_per_step": nq * 16,
                     "call": "eorb_matcher_search (host queries in, host records out; database resident)", "records_equal_device_path": same_h}
        except Exception as e:
            e2e_h = {"error": repr(e)}
    m.set_stream(None)
    api.nccl_comm_destroy(comm)
    parity, cpu_line = None, None
    if cpu and rank == 0:
        import oracle_lib as O
        cores = os.cpu_count() or 1
        db_host = np.empty((ndb_total, 32), np.uint8)
        for r in range(world):
            gr = torch.Generator(device="cuda"); gr.manual_seed(1234 + r)
            shard = d_db if r == rank else torch.randint(0, 256, (per, 32), dtype=torch.uint8, device="cuda", generator=gr)
            db_host[r * per:(r + 1) * per] = shard.cpu().numpy()
            del shard
        q_host = d_q.cpu().numpy()
        t0 = time.perf_counter()
        exp = O.hamming_best2(q_host, db_host, 50, 0.7, 0, cores)
        dt = time.perf_counter() - t0
        same = {k2: bool(np.array_equal(res[k1], exp[k2])) for k1, k2 in (("d", "best_dist"), ("i", "best_idx"), ("s", "second_dist"), ("a", "accepted"))}
        parity = {"records": int(nq), "database_rows": int(ndb_total), "shards": world, "fields_equal": same, "all_equal": all(same.values()),
                  "accepted": int(exp["accepted"].sum()), "checker": "oracle threaded popcount scan (lowest index wins ties), %d threads" % cores}
        cpu_line = {"value": pairs / dt / 1e9, "unit": "Gmatch/s", "cores": cores, "kind": "port",
                    "sample": "all %d queries x %d rows, __builtin_popcountll, cache-blocked scan, std::thread pool over query blocks, %.1f s" % (nq, ndb_total, dt)}
        del db_host
    bf16_peak = None
    try:
        bf16_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    tops = 2.0 * 256 * pairs / world / (ms * 1e-3) / 1e12
    return {"parity_checked": parity, "cpu_baseline": cpu_line, "e2e": e2e_h, "metric": "hamming_gmatch_per_s", "value": pairs / (ms * 1e-3) / 1e9, "unit": "Gmatch/s", "ms_per_step": ms,
            "workload": "configs[3]: 2000 queries x 16Mi rows, %d shard(s), best-2 + ratio 0.7; eorb_matcher_search_sharded (scan + ncclAllGather + merge)" % world,
            "engine": "tensor (tcgen05.mma kind::i8, exact +-1 / 0 contraction, dist = popc(q) - dot)" if engine_used == 1 else "popc",
            "engines_bit_identical": bool(bytes_popc == bytes_main),
            "planted_matches_found": ok, "gpu_launches": nlaunch,
            "roofline": {"bound": "tensor", "achieved": tops, "peak": (2.0 * bf16_peak) if bf16_peak else None, "unit": "TOP/s int8 per GPU",
                         "frac": (tops / (2.0 * bf16_peak)) if bf16_peak else None,
                         "note": "algorithmic 2 x 256 int8 ops per pair; peak = 2 x the measured dense bf16 rate of MEASURED_PEAKS.json (int8 runs at "
                                 "twice the bf16 rate). K = 256 is too short for the tensor pipe to be the limit: reading the int32 accumulator "
                                 "back from TMEM (4 B per pair) costs as much as the math, and both use the TMEM port (profiles/r02_hamming_tensor.md)"},
            "popc_engine": {"value": pairs / (ms_popc * 1e-3) / 1e9, "unit": "Gmatch/s", "ms_per_step": ms_popc,
                            "roofline": {"bound": "int-pipe (POPC)", "achieved": 8 * pairs / world / (ms_popc * 1e-3) / 1e12, "peak": popc_peak / 1e12,
                                         "unit": "TPOPC32/s per GPU", "frac": 8 * pairs / world / (ms_popc * 1e-3) / popc_peak,
                                         "note": "algorithmic 8 POPC32 per pair; peak measured live by eorb_probe_popc_rate"}}}


def bench_mci_jac(api, dev, steps, warmup):
    """one ev2mci_gg_f_jac call (the optimiser calls it once per iteration) on a 6000-event MC window."""
    import oracle_lib as O
    from eorb_slam_b200 import synth
    n, w, h = 6000, 240, 180
    ev = synth.make_events(n, seed=5, w=w, h=h)
    K = (199.09, 198.83, 132.19, 110.71)
    dt = float(ev["ts"][-1] - ev["ts"][0])
    T = synth.rotation_tcw(np.array([0.5, -0.7, 1.5]) * dt).astype(np.float64)
    R, t = T[:3, :3], np.array([0.02, -0.01, 0.03])
    cv = api.EvImConverter(dev, 1, n, w, h)
    for _ in range(max(warmup, 3)):
        got = cv.ev2mci_gg_f_jac(ev, K, R, t, 1.0, w, h, 1.0)
    reps = max(steps, 3) * 20
    t0 = time.perf_counter()
    for _ in range(reps):
        got = cv.ev2mci_gg_f_jac(ev, K, R, t, 1.0, w, h, 1.0)
    ms = (time.perf_counter() - t0) * 1e3 / reps
    t0 = time.perf_counter()
    for _ in range(5):
        exp = O.ev_mci_jac(ev, w, h, 1.0, R, t, 1.0, K)
    ms_port = (time.perf_counter() - t0) * 1e3 / 5
    return {"metric": "mci_jacobian_mev_per_s", "value": n / ms / 1e3, "unit": "Mev/s", "ms_per_call": ms,
            "workload": "ev2mci_gg_f_jac: %d events, 240x180, sigma 1: 7 splat frames (343 reductions per event) + 6 product means, host call" % n,
            "max_rel_err_vs_oracle": float(np.abs(got - exp).max() / np.abs(exp).max()),
            "cpu_baseline": {"kind": "port", "cores": 1, "ms_per_call": ms_port, "value": n / ms_port / 1e3, "unit": "Mev/s"}}


def bench_lk(api, torch, dev, steps, warmup):
    """ELK_Tracker on DAVIS240-shaped event frames, EvETHZ.yaml values (400 points, win 23, maxLevel 1,
    10 iterations, eps 0.03).  A step = one trackCurrImage call through the host C ABI (frame H2D, pyramid, tracker,
    points/status/err D2H).  CPU beside it: the oracle port and real OpenCV (cv2.calcOpticalFlowPyrLK) on the same input."""
    import oracle_lib as O
    from eorb_slam_b200 import synth
    per, w, h = 2000, 240, 180
    ev = synth.make_events(per * 2, seed=7, w=w, h=h)
    i0 = O.normalize_minmax_u8(O.ev_accumulate(ev[:per], w, h, 1.0, mode=1)[0])
    i1 = O.normalize_minmax_u8(O.ev_accumulate(ev[per // 2:per + per // 2], w, h, 1.0, mode=1)[0])
    _, kps, _ = O.OrbOracle(400, 1.0, 1, 0, 0, 9, w, h).extract(i0, (0, 1000), False)
    pts = np.stack([kps["x"], kps["y"]], 1).astype(np.float32)
    tr = api.ELK_Tracker(23, 1, 10, 0.03, dev, (w, h), len(pts))
    tr.setRefImage(i0, pts)
    for _ in range(max(warmup, 3)):
        tr.trackCurrImage(i1)
    n = max(steps, 3) * 20
    l0 = tr.launch_count()
    t0 = time.perf_counter()
    for _ in range(n):
        p, s, e = tr.trackCurrImage(i1)
    ms = (time.perf_counter() - t0) * 1e3 / n
    launches = (tr.launch_count() - l0) // n
    ep, es, ee, _ = O.lk_track(i0, i1, pts)
    t0 = time.perf_counter()
    for _ in range(5):
        O.lk_track(i0, i1, pts)
    ms_port = (time.perf_counter() - t0) * 1e3 / 5
    out = {"metric": "lk_tracked_points_per_s", "value": len(pts) / (ms * 1e-3), "unit": "points/s", "ms_per_call": ms,
           "workload": "ELK_Tracker: %d keypoints of a 240x180 event frame tracked into the next one, win 23, maxLevel 1, 10 it, eps 0.03" % len(pts),
           "gpu_launches_per_call": int(launches), "bit_exact_vs_oracle": bool(p.tobytes() == ep.tobytes() and np.array_equal(s, es)),
           "cpu_baseline": {"kind": "port", "cores": 1, "ms_per_call": ms_port, "value": len(pts) / (ms_port * 1e-3), "unit": "points/s"}}
    # the tracker with its state on the device: trackAndMatchCurrImage (LK from the resident last tracked points + refineTrackedPts on the
    # device, one D2H), against the plain call followed by the host bookkeeping it replaces
    tr2 = api.ELK_Tracker(23, 1, 10, 0.03, dev, (w, h), len(pts))
    tr2.setRefImageKPts(i0, kps)
    for _ in range(max(warmup, 3)):
        tr2.trackAndMatchCurrImage(i1)
    tr2.setRefImageKPts(i0, kps)
    t0 = time.perf_counter()
    for _ in range(n):
        nm, tk, m12, cnt, disp = tr2.trackAndMatchCurrImage(i1)
    ms_tm = (time.perf_counter() - t0) * 1e3 / n
    tr2.setRefImageKPts(i0, kps)
    nm, tk, m12, cnt, disp = tr2.trackAndMatchCurrImage(i1)
    enm, etk, em12, ecnt, edisp = O.lk_refine(ep, es, kps, w, h)
    out["track_and_match"] = {"ms_per_call": ms_tm, "matches": int(nm),
                              "bit_exact_vs_oracle": bool(nm == enm and tk.tobytes() == etk.tobytes() and np.array_equal(m12, em12)
                                                          and disp.tobytes() == edisp.tobytes()),
                              "note": "trackAndMatchCurrImage: tracked points stay in HBM as the next initial flow, refineTrackedPts on the device"}
    # the event front end of the L1 tracker with nothing but two counts leaving the device per window: events (HBM) -> event frame
    # (ev2im_gauss, normalised u8) -> trackAndMatchCurrImage against the reference frame's keypoints (EvAsynchTracker.cpp:588-594)
    try:
        nw = 16
        evs = synth.make_events(per * (nw + 1), seed=11, w=w, h=h)
        cvd = api.EvImConverter(dev, nw + 1, per, w, h)
        s2 = torch.cuda.Stream(); st2 = s2.cuda_stream
        cvd.set_stream(st2)
        tr3 = api.ELK_Tracker(23, 1, 10, 0.03, dev, (w, h), len(pts)); tr3.set_stream(st2)
        with torch.cuda.stream(s2):
            d_ev = torch.from_numpy(evs.view(np.uint8).reshape(-1)).cuda()
            d_f32 = torch.empty(h * w, dtype=torch.float32, device="cuda"); d_u8 = torch.empty(h * w, dtype=torch.uint8, device="cuda")
            d_tr = torch.zeros(len(pts) * 28, dtype=torch.uint8, device="cuda"); d_m = torch.zeros(len(pts), dtype=torch.uint8, device="cuda")
            d_dp = torch.zeros(len(pts), dtype=torch.float32, device="cuda"); d_c = torch.zeros(2, dtype=torch.int32, device="cuda")
            h_c = torch.zeros(2, dtype=torch.int32).pin_memory()
        pp = cvd.make_params(api.EV_GAUSS, w, h, 1.0, False, api.NORM_RUNNING)
        tr3.setRefImageKPts(i0, kps)

        def window(k):
            offs = np.array([k * per, (k + 1) * per], np.int64)
            cvd.accumulate_batch_device(d_ev.data_ptr(), offs, pp, d_f32.data_ptr(), d_u8.data_ptr())
            tr3.trackAndMatchCurrImage_device(d_u8.data_ptr(), w, d_tr.data_ptr(), d_m.data_ptr(), d_dp.data_ptr(), d_c.data_ptr())
            with torch.cuda.stream(s2):
                h_c.copy_(d_c, non_blocking=True)
            s2.synchronize()
            return int(h_c[0])
        for k in range(3):
            window(k)
        tr3.setRefImageKPts(i0, kps)
        t0 = time.perf_counter()
        nms = [window(k) for k in range(1, nw + 1)]
        ms_w = (time.perf_counter() - t0) * 1e3 / nw
        out["event_front_end"] = {"ms_per_window": ms_w, "matches_first_last": [nms[0], nms[-1]],
                                  "workload": "%d consecutive windows of %d events: events in HBM -> 240x180 event frame (u8) -> LK from the resident last "
                                              "tracked points + refineTrackedPts on the device; two counts come back per window" % (nw, per)}
        cvd.set_stream(None); tr3.set_stream(None)
    except Exception as ex:
        out["event_front_end"] = {"error": repr(ex)}
    try:
        import cv2
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 10, 0.03)
        t0 = time.perf_counter()
        for _ in range(20):
            cv2.calcOpticalFlowPyrLK(i0, i1, pts.reshape(-1, 1, 2), None, winSize=(23, 23), maxLevel=1, criteria=crit)
        ms_cv = (time.perf_counter() - t0) * 1e3 / 20
        out["opencv_cpu"] = {"ms_per_call": ms_cv, "value": len(pts) / (ms_cv * 1e-3), "unit": "points/s", "threads": cv2.getNumThreads(),
                             "note": "cv2 %s calcOpticalFlowPyrLK, the library the reference calls" % cv2.__version__}
    except Exception as ex:   # cv2 is optional
        out["opencv_cpu"] = {"error": repr(ex)}
    return out


def bench_guided(api, torch, dev, steps, warmup):
    """ORBmatcher::SearchForInitialization between two frames of the initialisation extractor
    (5 x nFeatures keypoints, 100-px window, ratio 0.9, rotation check).  Two numbers: the host call (keypoints +
    descriptors H2D, three kernels, matches D2H) and the device-resident call; CPU beside it: the oracle port."""
    import oracle_lib as O
    from eorb_slam_b200 import synth
    k1, d1, k2, d2, b = synth.make_keypoint_frame_pair(5000, 5000, 31)
    prev = np.stack([k1["x"], k1["y"]], 1)
    gm = api.GuidedMatcher(dev, 0.9, True)
    for _ in range(max(warmup, 3)):
        n, m12, p = gm.SearchForInitialization(k1, d1, k2, d2, b, prev, 100)
    reps = max(steps, 3) * 20
    l0 = gm.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        n, m12, p = gm.SearchForInitialization(k1, d1, k2, d2, b, prev, 100)
    ms = (time.perf_counter() - t0) * 1e3 / reps
    launches = (gm.launch_count() - l0) // reps
    # device-resident: inputs already in HBM, CUDA-event time of the three kernels
    tk1 = torch.from_numpy(k1.view(np.uint8).reshape(-1).copy()).cuda(); tk2 = torch.from_numpy(k2.view(np.uint8).reshape(-1).copy()).cuda()
    td1 = torch.from_numpy(d1).cuda(); td2 = torch.from_numpy(d2).cuda()
    tprev0 = torch.from_numpy(prev.astype(np.float32)).cuda(); tprev = tprev0.clone(); tm12 = torch.zeros(len(k1), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gm.set_stream(st)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    dev_ms = []
    for it in range(reps + 3):
        tprev.copy_(tprev0)
        e0.record()
        nd = gm.SearchForInitialization_device(tk1.data_ptr(), td1.data_ptr(), len(k1), tk2.data_ptr(), td2.data_ptr(), len(k2), b,
                                               tprev.data_ptr(), tm12.data_ptr(), 100)
        e1.record(); torch.cuda.synchronize()
        if it >= 3:
            dev_ms.append(e0.elapsed_time(e1))
    gm.set_stream(None)
    en, em12, ep = O.search_for_initialization(k1, d1, k2, d2, b, prev, 100, 0.9, True)
    t0 = time.perf_counter()
    for _ in range(5):
        O.search_for_initialization(k1, d1, k2, d2, b, prev, 100, 0.9, True)
    ms_port = (time.perf_counter() - t0) * 1e3 / 5
    lvl0 = int((k1["octave"] == 0).sum())
    # the tracking-rate caller: SearchByProjection(CurrentFrame, LastFrame, th = 15, bMono) between two 1000-feature frames
    proj = {}
    try:
        c = synth.make_projection_case(1009, 1009, 41)
        gp = api.GuidedMatcher(dev, 0.9, True)
        a = (c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], c["K"], c["scale_factors"])
        for _ in range(3):
            pn, pmc = gp.SearchByProjection(*a, 15.0)
        t0 = time.perf_counter()
        for _ in range(reps):
            pn, pmc = gp.SearchByProjection(*a, 15.0)
        pms = (time.perf_counter() - t0) * 1e3 / reps
        pen, pemc = O.search_by_projection(*a, 15.0, True)
        t0 = time.perf_counter()
        for _ in range(20):
            O.search_by_projection(*a, 15.0, True)
        pms_port = (time.perf_counter() - t0) * 1e3 / 20
        proj = {"ms_per_call": pms, "workload": "ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, 15, mono): 1009 x 1009 keypoints, %d valid map points, "
                "%d matches; host call" % (int(c["valid1"].sum()), pen), "bit_exact_vs_oracle": bool(pn == pen and np.array_equal(pmc, pemc)),
                "cpu_port_ms_per_call": pms_port}
    except Exception as e:
        proj = {"error": repr(e)}
    # Tracking::SearchLocalPoints: SearchByProjection(F, vpMapPoints, th = 3, ...) of a 3000-point local map against one 1000-feature frame
    loc = {}
    try:
        c = synth.make_local_map_case(3000, 1009, 43)
        gl = api.GuidedMatcher(dev, 0.8, True)
        a = (c["pts"], c["descMP"], c["kps2"], c["desc2"], c["held2"], c["bounds"], c["scale_factors"])
        for _ in range(3):
            ln, lmc = gl.SearchByProjectionMapPoints(*a, 3.0, False, 0.0)
        t0 = time.perf_counter()
        for _ in range(reps):
            ln, lmc = gl.SearchByProjectionMapPoints(*a, 3.0, False, 0.0)
        lms = (time.perf_counter() - t0) * 1e3 / reps
        len_, lemc = O.search_by_projection_map_points(*a, 3.0, False, 0.0, 0.8)
        t0 = time.perf_counter()
        for _ in range(20):
            O.search_by_projection_map_points(*a, 3.0, False, 0.0, 0.8)
        lms_port = (time.perf_counter() - t0) * 1e3 / 20
        loc = {"ms_per_call": lms, "workload": "ORBmatcher::SearchByProjection(F, vpMapPoints, 3): 3000 local map points (%d in view) x 1009 keypoints, "
               "%d matches; host call" % (int(c["pts"]["in_view"].sum()), len_), "bit_exact_vs_oracle": bool(ln == len_ and np.array_equal(lmc, lemc)),
               "cpu_port_ms_per_call": lms_port}
    except Exception as e:
        loc = {"error": repr(e)}
    # LocalMapping::SearchInNeighbors / LoopClosing::SearchAndFuse: ORBmatcher::Fuse's matching core, 3000 map points into a 1009-feature keyframe
    # (windows formed by the caller as the reference's host code does, reprojection gate on the device, no blocking)
    fus = {}
    try:
        c = synth.make_local_map_case(3000, 1009, 47)
        q = np.zeros(3000, api.AREA_QUERY_DTYPE)
        lv = np.clip(c["pts"]["scale_level"], 0, len(c["scale_factors"]) - 1)
        q["x"] = c["pts"]["proj_x"]; q["y"] = c["pts"]["proj_y"]; q["r"] = np.float32(3.0) * c["scale_factors"][lv]
        q["min_level"] = lv - 1; q["max_level"] = lv
        inv = (np.float32(1.0) / (c["scale_factors"] ** 2)).astype(np.float32)
        gf = api.GuidedMatcher(dev, 0.8, True)
        a = (q, None, c["descMP"], c["kps2"], c["desc2"], None, None, c["bounds"])
        kw = dict(inv_level_sigma2=inv, blocking=False, th_high=50)
        for _ in range(3):
            fr = gf.SearchWindows(*a, **kw)
        t0 = time.perf_counter()
        for _ in range(reps):
            fr = gf.SearchWindows(*a, **kw)
        fms = (time.perf_counter() - t0) * 1e3 / reps
        fo = O.search_windows(*a, **kw)
        t0 = time.perf_counter()
        for _ in range(20):
            O.search_windows(*a, **kw)
        fms_port = (time.perf_counter() - t0) * 1e3 / 20
        fus = {"ms_per_call": fms, "workload": "matching core of ORBmatcher::Fuse (eorb_guided_search_windows): 3000 map points x 1009 keypoints, th 3, reprojection "
               "gate, %d fusions; host call" % fo[0], "bit_exact_vs_oracle": bool(fr[0] == fo[0] and all(np.array_equal(x, y) for x, y in zip(fr[1:], fo[1:]))),
               "cpu_port_ms_per_call": fms_port}
    except Exception as e:
        fus = {"error": repr(e)}
    return {"search_by_projection": proj, "search_local_points": loc, "fuse_core": fus, "metric": "search_for_initialization_calls_per_s", "value": 1e3 / ms, "unit": "calls/s", "ms_per_call": ms,
            "ms_per_call_device_resident": float(np.median(dev_ms)),
            "workload": "ORBmatcher::SearchForInitialization: 5000 x 5000 keypoints (%d level-0 queries), window 100, ratio 0.9, "
                        "rotation check; %d matches" % (lvl0, en),
            "gpu_launches_per_call": int(launches),
            "bit_exact_vs_oracle": bool(n == en and nd == en and np.array_equal(m12, em12) and p.tobytes() == ep.tobytes()
                                        and np.array_equal(tm12.cpu().numpy(), em12)),
            "cpu_baseline": {"kind": "port", "cores": 1, "ms_per_call": ms_port, "value": 1e3 / ms_port, "unit": "calls/s"}}


def bench_bow(api, torch, dev, steps, warmup):
    """Frame::ComputeBoW = ORBVocabulary::transform(descriptors, BowVector, FeatureVector, 4) for one frame's
    descriptors against a synthetic vocabulary of ORBvoc.txt's size (k = 10, L = 6), host call; plus Frame::UndistortKeyPoints."""
    import oracle_lib as O
    from eorb_slam_b200 import synth
    voc = synth.make_vocabulary_regular(10, 6, 3)
    leaves = np.flatnonzero(voc["is_leaf"])
    rng = np.random.default_rng(4)
    feats = voc["desc"][rng.choice(leaves, 1009)].copy()
    feats[:, :4] ^= rng.integers(0, 256, (1009, 4), dtype=np.uint8)          # perturbed words
    v = api.ORBVocabulary(voc, dev)
    for _ in range(max(warmup, 3)):
        got = v.transform(feats, 4)
    reps = max(steps, 3) * 20
    l0 = v.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        got = v.transform(feats, 4)
    ms = (time.perf_counter() - t0) * 1e3 / reps
    launches = (v.launch_count() - l0) // reps
    orc = O.VocabOracle(voc)
    exp = orc.transform(feats, 4)
    t0 = time.perf_counter()
    for _ in range(5):
        orc.transform(feats, 4)
    ms_port = (time.perf_counter() - t0) * 1e3 / 5
    same = all(np.array_equal(got[k], exp[k]) for k in ("bow_ids", "fv_nodes", "fv_start", "fv_feats")) and got["bow_vals"].tobytes() == exp["bow_vals"].tobytes()
    k1, _, _, _, _ = synth.make_keypoint_frame_pair(1009, 10, 13)
    K, D = (458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)
    api.UndistortKeyPoints(k1, K, D)
    t0 = time.perf_counter()
    for _ in range(reps):
        un = api.UndistortKeyPoints(k1, K, D)
    ms_un = (time.perf_counter() - t0) * 1e3 / reps
    # Tracking::TrackReferenceKeyFrame: ComputeBoW on the frame, then SearchByBoW(pKF, F, matches) with ORBmatcher(0.7, true)
    sbb = {}
    try:
        kk1, _, kk2, _, _ = synth.make_keypoint_frame_pair(1009, 1009, 17)
        dkf = feats
        df = feats.copy()
        flips = rng.integers(0, 256, (len(df), 2))
        df[np.arange(len(df)), flips[:, 0] % 28 + 4] ^= (1 << (flips[:, 1] % 8)).astype(np.uint8)     # the frame sees the keyframe's features, a few bits off
        order = rng.permutation(len(df)); df = df[order].copy(); kk2 = kk2[:len(df)]
        t1 = v.transform(dkf, 4); t2 = v.transform(df, 4)
        fv1 = (t1["fv_nodes"], t1["fv_start"], t1["fv_feats"]); fv2 = (t2["fv_nodes"], t2["fv_start"], t2["fv_feats"])
        valid = np.ones(len(dkf), np.uint8)
        gm = api.GuidedMatcher(dev, 0.7, False)
        for _ in range(3):
            bn, bmf = gm.SearchByBoW(kk1, dkf, valid, fv1, kk2, df, fv2)
        t0 = time.perf_counter()
        for _ in range(reps):
            bn, bmf = gm.SearchByBoW(kk1, dkf, valid, fv1, kk2, df, fv2)
        bms = (time.perf_counter() - t0) * 1e3 / reps
        ben, bemf = O.search_by_bow(kk1, dkf, valid, fv1, kk2, df, fv2, 0.7, False)
        t0 = time.perf_counter()
        for _ in range(20):
            O.search_by_bow(kk1, dkf, valid, fv1, kk2, df, fv2, 0.7, False)
        bms_port = (time.perf_counter() - t0) * 1e3 / 20
        sbb = {"ms_per_call": bms, "workload": "ORBmatcher::SearchByBoW(pKF, F): 1009 x 1009 features over %d / %d vocabulary nodes, %d matches; host call" %
               (len(fv1[0]), len(fv2[0]), ben), "bit_exact_vs_oracle": bool(bn == ben and np.array_equal(bmf, bemf)), "cpu_port_ms_per_call": bms_port}
    except Exception as e:
        sbb = {"error": repr(e)}
    # LocalMapping::CreateNewMapPoints: SearchForTriangulation between the same two feature sets (coarse = the descriptor search + epipole gate)
    tri = {}
    try:
        nl = 8
        sc = (np.float32(1.2) ** np.arange(nl)).astype(np.float32); sg = (sc * sc).astype(np.float32)
        F12 = np.array([0, 0, 0, 0, 0, -1e-4, 0, 1e-4, 0], np.float32); ep = np.array([-np.inf, 248.0], np.float32)   # sideways motion: horizontal epipolar lines
        f1 = O.triangulation_flags(np.zeros(len(kk1)), None, False); f2 = O.triangulation_flags(np.zeros(len(kk2)), None, False)
        a = (kk1, dkf, f1, fv1, kk2, df, f2, fv2, F12, ep, sc, sg)
        gt = api.GuidedMatcher(dev, 0.6, True)
        for _ in range(3):
            tn, tm = gt.SearchForTriangulation(*a, bCoarse=True)
        t0 = time.perf_counter()
        for _ in range(reps):
            tn, tm = gt.SearchForTriangulation(*a, bCoarse=True)
        tms = (time.perf_counter() - t0) * 1e3 / reps
        ten, tem = O.search_for_triangulation(*a, coarse=True, check_ori=True)
        t0 = time.perf_counter()
        for _ in range(20):
            O.search_for_triangulation(*a, coarse=True, check_ori=True)
        tms_port = (time.perf_counter() - t0) * 1e3 / 20
        tn2, tm2 = gt.SearchForTriangulation(*a, bCoarse=False)
        ten2, tem2 = O.search_for_triangulation(*a, coarse=False, check_ori=True)
        tri = {"ms_per_call": tms, "workload": "ORBmatcher::SearchForTriangulation(pKF1, pKF2, bCoarse): 1009 x 1009 features over %d / %d vocabulary nodes, %d pairs; "
               "host call" % (len(fv1[0]), len(fv2[0]), ten),
               "bit_exact_vs_oracle": bool(tn == ten and np.array_equal(tm, tem) and tn2 == ten2 and np.array_equal(tm2, tem2)), "cpu_port_ms_per_call": tms_port}
    except Exception as e:
        tri = {"error": repr(e)}
    return {"search_by_bow": sbb, "search_for_triangulation": tri, "metric": "bow_transform_features_per_s", "value": len(feats) / (ms * 1e-3), "unit": "features/s", "ms_per_call": ms,
            "workload": "ORBVocabulary::transform: %d descriptors, vocabulary k=10 L=6 (%d nodes, %d words), levelsup 4, host call" %
                        (len(feats), len(voc["parent"]), len(leaves)),
            "gpu_launches_per_call": int(launches), "bit_exact_vs_oracle": bool(same), "bow_words": int(len(exp["bow_ids"])),
            "cpu_baseline": {"kind": "port", "cores": 1, "ms_per_call": ms_port, "value": len(feats) / (ms_port * 1e-3), "unit": "features/s"},
            "undistort_keypoints": {"ms_per_call": ms_un, "n": len(k1), "note": "Frame::UndistortKeyPoints, host call (malloc + H2D + kernel + D2H)"}}


def bench_chain(api, torch, dev, steps, warmup):
    """The tracking thread's per-frame chain with the frame data resident in HBM (Tracking::TrackReferenceKeyFrame + SearchLocalPoints,
    src/Tracking-1.cc:1690, 1818, 2436): frame (host) -> ORB extraction -> Frame::UndistortKeyPoints -> Frame::ComputeBoW ->
    ORBmatcher(0.7).SearchByBoW(pKF, F) -> ORBmatcher(0.8).SearchByProjection(F, local map points, th = 3).  Only the frame, three counts
    and the two match tables cross PCIe; keypoints, descriptors and the FeatureVector never leave the device.  Next to it: the same
    five steps through the CPU oracle ports, one core, and the result check."""
    import oracle_lib as O
    from eorb_slam_b200 import synth
    Wc, Hc = 752, 480
    img1 = synth.make_frame(31)
    img2 = np.roll(img1, (2, -3), axis=(0, 1))
    K, D = (458.654, 457.296, 367.215, 248.375), (0.0, 0.0, 0.0, 0.0, 0.0)   # rectified input: UndistortKeyPoints copies (Frame.cc:807-811)
    voc = synth.make_vocabulary(10, 4, 7)
    p = api.ORBxParams(1000, 1.2, 8, 20, 7, 19, (Wc, Hc))
    ex = api.ORBextractor(p, dev, 1)
    cap = ex.cap
    s = torch.cuda.Stream()
    st = s.cuda_stream
    ex.set_stream(st)
    v = api.ORBVocabulary(voc, dev); v.set_stream(st)
    gm = api.GuidedMatcher(dev, 0.7, True); gm.set_stream(st)
    gl = api.GuidedMatcher(dev, 0.8, True); gl.set_stream(st)
    sf = np.ones(8, np.float32)
    for i in range(1, 8):
        sf[i] = np.float32(np.float64(sf[i - 1]) * np.float64(np.float32(1.2)))
    bnd = np.array([0, 0, Wc, Hc], np.float32)
    with torch.cuda.stream(s):
        h_img = torch.from_numpy(np.stack([img1, img2])).pin_memory()
        d_img = torch.empty((Hc, Wc), dtype=torch.uint8, device="cuda")
        d_kps = [torch.zeros(cap * 28, dtype=torch.uint8, device="cuda") for _ in range(2)]
        d_un = torch.zeros(cap * 28, dtype=torch.uint8, device="cuda")
        d_desc = [torch.zeros(cap * 32, dtype=torch.uint8, device="cuda") for _ in range(2)]
        d_n = torch.zeros(1, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(1, dtype=torch.int32, device="cuda")
        h_n = torch.zeros(1, dtype=torch.int32).pin_memory()
        fv = [[torch.zeros(cap + 1, dtype=torch.int32, device="cuda") for _ in range(3)] for _ in range(2)]
        d_mf = torch.zeros(cap, dtype=torch.int32, device="cuda"); d_mc = torch.zeros(cap, dtype=torch.int32, device="cuda")
        h_mf = torch.zeros(cap, dtype=torch.int32).pin_memory(); h_mc = torch.zeros(cap, dtype=torch.int32).pin_memory()

        def extract(k):
            d_img.copy_(h_img[k], non_blocking=True)
            ex.extract_batch_raw(d_img.data_ptr(), 1, Wc, Hc, Wc, Wc * Hc, (0, 0), True, d_kps[k].data_ptr(), d_desc[k].data_ptr(), cap,
                                 d_n.data_ptr(), d_mono.data_ptr(), device=True)
            h_n.copy_(d_n, non_blocking=True)
            s.synchronize()
            return int(h_n[0])

        # the keyframe: extracted once, FeatureVector kept on the device, 85 % of its features carry a map point
        n1 = extract(0)
        _, nf1 = v.transform_resident(d_desc[0].data_ptr(), n1, 2, fv[0][0].data_ptr(), fv[0][1].data_ptr(), fv[0][2].data_ptr())
        rng = np.random.default_rng(5)
        valid = (rng.random(n1) < 0.85).astype(np.uint8)
        d_valid = torch.from_numpy(valid).cuda()
        s.synchronize()
        k1 = np.frombuffer(d_kps[0].cpu().numpy().tobytes(), api.KEYPOINT_DTYPE)[:n1]
        pts = np.zeros(n1, synth.TRACK_POINT_DTYPE)
        pts["proj_x"] = k1["x"] - 3.0 + rng.normal(0, 0.7, n1).astype(np.float32); pts["proj_y"] = k1["y"] + 2.0 + rng.normal(0, 0.7, n1).astype(np.float32)
        pts["view_cos"] = rng.uniform(0.99, 1.0, n1).astype(np.float32); pts["depth"] = 5.0
        pts["scale_level"] = k1["octave"]; pts["observations"] = rng.integers(1, 4, n1); pts["in_view"] = 1
        d_pts = torch.from_numpy(pts.view(np.uint8).copy()).cuda()
        d_held = torch.zeros(cap, dtype=torch.uint8, device="cuda")

        def frame_step():
            n2 = extract(1)
            api.UndistortKeyPoints_device(d_kps[1].data_ptr(), d_un.data_ptr(), n2, K, D, st)
            _, nf2 = v.transform_resident(d_desc[1].data_ptr(), n2, 2, fv[1][0].data_ptr(), fv[1][1].data_ptr(), fv[1][2].data_ptr())
            nb = gm.SearchByBoW_device(d_kps[0].data_ptr(), d_desc[0].data_ptr(), d_valid.data_ptr(), n1, tuple(t.data_ptr() for t in fv[0]), nf1,
                                       d_un.data_ptr(), d_desc[1].data_ptr(), n2, tuple(t.data_ptr() for t in fv[1]), nf2, d_mf.data_ptr())
            torch.ge(d_mf[:n2], 0, out=d_held[:n2].view(torch.bool))          # slots TrackReferenceKeyFrame filled are held in the local-map search
            nl = gl.SearchByProjectionMapPoints_device(d_pts.data_ptr(), d_desc[0].data_ptr(), n1, d_un.data_ptr(), d_desc[1].data_ptr(),
                                                       d_held.data_ptr(), n2, bnd, sf, d_mc.data_ptr(), 3.0)
            h_mf[:n2].copy_(d_mf[:n2], non_blocking=True); h_mc[:n2].copy_(d_mc[:n2], non_blocking=True)
            s.synchronize()
            return n2, nb, nl

        for _ in range(max(warmup, 5)):
            n2, nb, nl = frame_step()
        reps = max(steps, 3) * 20
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); frame_step(); ts.append(time.perf_counter() - t0)
        ms = float(np.median(ts)) * 1e3
        mf = h_mf[:n2].numpy().copy(); mc = h_mc[:n2].numpy().copy()
    for hdl in (ex, v, gm, gl):
        hdl.set_stream(None)
    # the same chain through the CPU ports (one core) + the result check
    orc = O.OrbOracle(1000, 1.2, 8, 20, 7, 19, Wc, Hc)
    vo = O.VocabOracle(voc)
    _, ok1, od1 = orc.extract(img1, (0, 0), True)
    e1 = vo.transform(od1, 2)
    cpu = {}
    t0 = time.perf_counter(); _, ok2, od2 = orc.extract(img2, (0, 0), True); cpu["extract"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); e2 = vo.transform(od2, 2); cpu["transform"] = (time.perf_counter() - t0) * 1e3
    f1 = (e1["fv_nodes"], e1["fv_start"], e1["fv_feats"]); f2 = (e2["fv_nodes"], e2["fv_start"], e2["fv_feats"])
    t0 = time.perf_counter(); en, emf = O.search_by_bow(ok1, od1, valid, f1, ok2, od2, f2, 0.7, True); cpu["search_by_bow"] = (time.perf_counter() - t0) * 1e3
    held = (emf >= 0).astype(np.uint8)
    t0 = time.perf_counter()
    eln, elmc = O.search_by_projection_map_points(pts, od1, ok2, od2, held, bnd, sf, 3.0, False, 0.0, 0.8)
    cpu["search_local_points"] = (time.perf_counter() - t0) * 1e3
    ok = bool(n2 == len(ok2) and nb == en and np.array_equal(mf, emf) and nl == eln and np.array_equal(mc, elmc))
    return {"metric": "tracking_chain_latency_ms", "value": ms, "unit": "ms", "p90": float(np.percentile(ts, 90)) * 1e3,
            "workload": "one 752x480 frame (pinned host) -> extract -> undistort -> ComputeBoW (k=10 L=4 vocabulary, levelsup 2) -> SearchByBoW against a "
                        "%d-feature keyframe -> SearchByProjection(F, %d local map points, th 3); %d + %d matches; data resident in HBM, "
                        "3 counts + 2 match tables come back" % (n1, n1, nb, nl),
            "bit_exact_vs_oracle": ok,
            "cpu_baseline": {"kind": "port", "cores": 1, "ms_per_frame": float(sum(cpu.values())), "stages_ms": cpu}}


# ------------------------------------------------------------------------------------------------ main arm
def run_ours(args):
    import torch
    from eorb_slam_b200 import api, synth
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if api.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; eorb_slam_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = local
    numa = _bind_to_gpu_cpus(local)      # pinned host buffers are first-touched on the GPU's NUMA node
    nfr = args.frames
    chunk = min(args.chunk, nfr)

    frames = make_batch(nfr, seed0=rank * 100000)
    p = api.ORBxParams(ORB_KW["nfeatures"], ORB_KW["scale_factor"], ORB_KW["nlevels"], ORB_KW["ini_th"], ORB_KW["min_th"],
                       ORB_KW["edge_th"], (W, H))
    ex = api.ORBextractor(p, dev, chunk)
    cap = ex.cap
    st = torch.cuda.current_stream().cuda_stream
    ex.set_stream(st)

    # ---- resident inputs / outputs
    h_frames = torch.from_numpy(frames).pin_memory()
    d_frames = h_frames.cuda(non_blocking=True)
    d_kps = torch.empty(nfr * cap * 28, dtype=torch.uint8, device="cuda")
    d_desc = torch.empty(nfr * cap * 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(nfr, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(nfr, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    def step_device():
        for f0 in range(0, nfr, chunk):
            nb = min(chunk, nfr - f0)
            ex.extract_batch_raw(d_frames.data_ptr() + f0 * W * H, nb, W, H, W, W * H, (0, 1000), True,
                                 d_kps.data_ptr() + f0 * cap * 28, d_desc.data_ptr() + f0 * cap * 32, cap,
                                 d_n.data_ptr() + f0 * 4, d_mono.data_ptr() + f0 * 4, device=True)

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # headline: no stage events inside the timed region
    l0 = ex.launch_count()
    clocks = ClockSampler(local)
    clocks.start()
    t = api.CudaTimer()
    t.start(st)
    for _ in range(args.steps):
        step_device()
    t.stop(st)
    ms_total = t.elapsed_ms()
    torch.cuda.synchronize()
    clk = clocks.stop()
    launches = ex.launch_count() - l0
    # per-kernel times: the same steps again with an event pair around every stage
    ex.stage_timing(True)
    t.start(st)
    for _ in range(args.steps):
        step_device()
    t.stop(st)
    ms_serial = t.elapsed_ms() / args.steps
    stage_ms, stage_launches = ex.stage_times()
    ex.stage_timing(False)
    if world > 1:
        tm = torch.tensor([ms_total], device="cuda"); dist.all_reduce(tm, op=dist.ReduceOp.MAX); ms_total = float(tm.item())
        dist.barrier()
    ms_step = ms_total / args.steps
    value = world * nfr / (ms_step * 1e-3)
    nkp = int(d_n.sum().item())

    # ---- PCIe H2D rate of the same pinned batch (the ceiling of the end-to-end number: 360 960 B per frame must cross it)
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    d_frames.copy_(h_frames, non_blocking=True); torch.cuda.synchronize()
    ev0.record()
    for _ in range(2):
        d_frames.copy_(h_frames, non_blocking=True)
    ev1.record(); torch.cuda.synchronize()
    h2d_gbs = 2 * nfr * W * H / (ev0.elapsed_time(ev1) * 1e-3) / 1e9

    # ---- end to end: pinned host frames -> C-ABI host call -> pinned host keypoints/descriptors
    h_kps = torch.empty(nfr * cap * 28, dtype=torch.uint8).pin_memory()
    h_desc = torch.empty(nfr * cap * 32, dtype=torch.uint8).pin_memory()
    h_n = torch.zeros(nfr, dtype=torch.int32).pin_memory(); h_mono = torch.zeros(nfr, dtype=torch.int32).pin_memory()
    ex.set_stream(None)

    def step_e2e():
        ex.extract_batch_raw(h_frames.data_ptr(), nfr, W, H, W, W * H, (0, 1000), True, h_kps.data_ptr(), h_desc.data_ptr(), cap,
                             h_n.data_ptr(), h_mono.data_ptr(), device=False)

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    if world > 1:
        dist.barrier()
    e2e_steps = max(args.steps, 10)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        tm = torch.tensor([e2e_ms], device="cuda"); dist.all_reduce(tm, op=dist.ReduceOp.MAX); e2e_ms = float(tm.item())
    e2e_value = world * nfr / (e2e_ms * 1e-3)
    assert int(h_n.sum().item()) == nkp, "e2e path and resident path disagree"
    h2d = nfr * W * H
    d2h = nfr * (cap * 60 + 8)

    # ---- roofline of the dominant kernel
    peak, peak_src = _peaks()
    dom = max(stage_ms, key=stage_ms.get)
    frames_timed = nfr * args.steps
    per_stage = {}
    traffic, traffic_src = _traffic()
    for k, v in stage_ms.items():
        e = {"ms_per_frame": v / frames_timed, "share": v / max(sum(stage_ms.values()), 1e-9), "launches": stage_launches[k]}
        if k in ALGO_BYTES:
            e["algo_bytes_per_frame"] = ALGO_BYTES[k]
            e["achieved_gbs"] = ALGO_BYTES[k] * frames_timed / (v * 1e-3) / 1e9 if v > 0 else None
            e["frac_of_hbm_peak"] = e["achieved_gbs"] / peak if v > 0 else None
        if k in traffic:
            e["ncu_dram_bytes_per_frame"] = traffic[k]
        wi = getattr(_traffic, "inst", {}).get(k)
        if wi and v > 0:
            # instruction-issue roofline: executed warp instructions per frame (ncu, profiles/r02_traffic.json) over the live
            # CUDA-event time, against SMs x 4 schedulers x 1 warp instruction per clock at the sampled SM clock
            e["warp_inst_per_frame"] = wi
            e["issue_slots_frac"] = (wi * frames_timed / (v * 1e-3)) / (148 * 4 * (clk.get("sm_mhz") or 1965.0) * 1e6)
        per_stage[k] = e
    if dom in ALGO_BYTES:
        dom_bytes_launch = ALGO_BYTES[dom] * chunk / (stage_launches[dom] / (frames_timed / chunk))
        dom_ms_launch = stage_ms[dom] / stage_launches[dom]
        achieved = dom_bytes_launch / (dom_ms_launch * 1e-3) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic[dom] * chunk / (stage_launches[dom] / (frames_timed / chunk))) if dom in traffic else None,
                "traffic_source": ("profiles/r02_traffic.json (%s: ncu --set full of this command at %s frames per launch set), bytes per frame x frames per launch" % (traffic_src, getattr(_traffic, "frames_per_launch", None))) if dom in traffic else None,
                "peak_source": peak_src, "avg_launch_ms": dom_ms_launch, "algo_bytes_per_launch": dom_bytes_launch,
                "note": "byte-granular integer kernel: issue-bound on the ALU pipe, not on HBM (see profiles/ and DESIGN.md)",
                "issue_roofline": {"bound": "instruction issue (integer pipes)", "frac": per_stage[dom].get("issue_slots_frac"),
                                   "warp_inst_per_frame": per_stage[dom].get("warp_inst_per_frame"),
                                   "peak": "148 SMs x 4 schedulers x 1 warp instruction / clk"}}
    else:
        # latency-bound stage (octree / index / orient+desc): no HBM roofline applies; report the best HBM-bound stage too
        hb = max((k for k in ALGO_BYTES), key=lambda k: stage_ms[k])
        achieved = ALGO_BYTES[hb] * frames_timed / (stage_ms[hb] * 1e-3) / 1e9
        roof = {"kernel": hb, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic[hb] * chunk / (stage_launches[hb] / (frames_timed / chunk))) if hb in traffic else None,
                "peak_source": peak_src, "avg_launch_ms": stage_ms[hb] / max(stage_launches[hb], 1),
                "note": "largest share of the step is '%s' (latency-bound, no HBM roofline); this entry is the largest HBM-bound kernel" % dom}

    extra = {}
    if not args.no_extras:
        try:
            extra["events"] = bench_events(api, torch, dev, max(args.steps, 20), max(args.warmup, 3), world, rank, dist, not args.no_cpu)   # 0.07 ms steps: 20 for a stable mean
        except Exception as e:   # extras never invalidate the headline line
            extra["events"] = {"error": repr(e)}
        try:
            if rank == 0:   # configs[0]: ONE frame through ORBextractor::operator() (host image in, keypoints + descriptors out)
                from eorb_slam_b200 import synth as _s
                one = api.ORBextractor(p, dev, 1)
                fr = [_s.make_frame(900 + i) for i in range(4)]
                for i in range(10):
                    one(fr[i % 4])
                lat = []
                for i in range(100):
                    t0 = time.perf_counter(); one(fr[i % 4]); lat.append((time.perf_counter() - t0) * 1e3)
                extra["single_frame"] = {"metric": "orb_single_frame_latency_ms", "value": float(np.median(lat)), "unit": "ms",
                                         "p90": float(np.percentile(lat, 90)),
                                         "workload": "configs[0]: one 752x480 frame per call, pageable host image, 12 kernels, results on the host"}
        except Exception as e:
            extra["single_frame"] = {"error": repr(e)}
        try:
            if rank == 0:
                extra["lk"] = bench_lk(api, torch, dev, args.steps, args.warmup)
        except Exception as e:
            extra["lk"] = {"error": repr(e)}
        try:
            if rank == 0:
                extra["mci_jac"] = bench_mci_jac(api, dev, args.steps, args.warmup)
        except Exception as e:
            extra["mci_jac"] = {"error": repr(e)}
        try:
            if rank == 0:
                extra["guided"] = bench_guided(api, torch, dev, args.steps, args.warmup)
        except Exception as e:
            extra["guided"] = {"error": repr(e)}
        try:
            if rank == 0:
                extra["bow"] = bench_bow(api, torch, dev, args.steps, args.warmup)
        except Exception as e:
            extra["bow"] = {"error": repr(e)}
        try:
            if rank == 0:
                extra["tracking_chain"] = bench_chain(api, torch, dev, args.steps, args.warmup)
        except Exception as e:
            extra["tracking_chain"] = {"error": repr(e)}
        try:
            extra["hamming"] = bench_hamming(api, torch, dev, max(min(args.steps, 3), 1), 3, world, rank, dist, not args.no_cpu)
        except Exception as e:
            extra["hamming"] = {"error": repr(e)}

    cpu, cpu_cv2, parity = None, None, None
    if rank == 0 and not args.no_cpu:
        # ---- parity of the WHOLE timed batch, outside the timed region: every frame's keypoint count against the oracle (the
        #      pinned restatement: == the reference's own code, tests/test_ref_pin.py) and the first frames' full keypoint and
        #      descriptor bytes.  The oracle pass over the batch is CPU work on all host cores.
        import oracle_lib as O
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        _, ocounts = O.orb_extract_batch_mt(frames, cores)
        dt_orc = time.perf_counter() - t0
        gcounts = d_n.cpu().numpy()
        nfull = min(nfr, 64)
        gk = d_kps[:nfull * cap * 28].cpu().numpy().reshape(nfull, cap * 28)
        gd = d_desc[:nfull * cap * 32].cpu().numpy().reshape(nfull, cap, 32)
        orc = O.OrbOracle(ORB_KW["nfeatures"], ORB_KW["scale_factor"], ORB_KW["nlevels"], ORB_KW["ini_th"], ORB_KW["min_th"], ORB_KW["edge_th"], W, H)
        full_ok = 0
        for f in range(nfull):
            _, ok_, od_ = orc.extract(frames[f])
            n_ = len(ok_)
            full_ok += int(n_ == int(gcounts[f]) and gk[f][:n_ * 28].tobytes() == ok_.tobytes() and np.array_equal(gd[f][:n_], od_))
        parity = {"frames_count_checked": int(nfr), "frame_counts_equal": int((gcounts == ocounts).sum()),
                  "frames_full_checked": int(nfull), "frames_full_equal": int(full_ok),
                  "all_equal": bool((gcounts == ocounts).all() and full_ok == nfull),
                  "checker": "oracle (pinned byte for byte to the reference's ORBextractor.cc, tests/golden/ref_*.npz), %d threads, %.1f s" % (cores, dt_orc)}
        if world == 1:
            fn, kind, what = _cpu_reference_extractor()
            fn(frames[:cores], cores)
            t0 = time.perf_counter()
            fn(frames, cores)
            dt = time.perf_counter() - t0
            cpu = {"value": nfr / dt, "unit": "frames/s", "cores": cores, "kind": kind,
                   "sample": "all %d frames of the same batch, std::thread pool over frames, one extractor per thread, %.1f s; %s" % (nfr, dt, what),
                   "oracle_port_frames_per_s": nfr / dt_orc}
            try:    # BASELINE.md section 3 line 2, in its own process (a fork pool beside an initialised CUDA context is not safe)
                import subprocess
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cv2-baseline"], capture_output=True, text=True, timeout=600)
                cpu_cv2 = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as e:
                cpu_cv2 = {"error": repr(e)}

    # ---- strong scaling of configs[2] as written: 4096 frames in total, split by frame over the ranks
    strong = None
    if world > 1:
        share = FRAMES_PER_GPU // world

        def step_share():
            for f0 in range(0, share, chunk):
                nb = min(chunk, share - f0)
                ex.extract_batch_raw(d_frames.data_ptr() + f0 * W * H, nb, W, H, W, W * H, (0, 1000), True,
                                     d_kps.data_ptr() + f0 * cap * 28, d_desc.data_ptr() + f0 * cap * 32, cap,
                                     d_n.data_ptr() + f0 * 4, d_mono.data_ptr() + f0 * 4, device=True)
        ex.set_stream(st)
        for _ in range(3):
            step_share()
        torch.cuda.synchronize(); dist.barrier()
        t.start(st)
        for _ in range(args.steps):
            step_share()
        t.stop(st)
        ms_s = t.elapsed_ms() / args.steps
        tm = torch.tensor([ms_s], device="cuda"); dist.all_reduce(tm, op=dist.ReduceOp.MAX); ms_s = float(tm.item())
        strong = {"metric": "orb_frames_per_sec", "scaling": "strong", "value": share * world / (ms_s * 1e-3), "unit": "frames/s",
                  "ms_per_step": ms_s, "frames_total": share * world, "frames_per_gpu": share,
                  "workload": "configs[2] as written: %d frames in total, partitioned by frame over %d GPUs, no collective" % (share * world, world)}
        ex.set_stream(None)

    # ---- host link: PCIe generation / width per GPU, and the H2D rate of every GPU alone and with all ranks copying at once
    link = None
    try:
        import pynvml
        pynvml.nvmlInit()
        hh = pynvml.nvmlDeviceGetHandleByIndex(local)
        mine = [float(pynvml.nvmlDeviceGetCurrPcieLinkGeneration(hh)), float(pynvml.nvmlDeviceGetCurrPcieLinkWidth(hh))]
    except Exception:
        mine = [0.0, 0.0]
    if world > 1:
        def h2d_rate():
            ev0.record()
            d_frames.copy_(h_frames, non_blocking=True)
            ev1.record(); torch.cuda.synchronize()
            return nfr * W * H / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
        solo = 0.0
        for r in range(world):
            dist.barrier()
            if r == rank:
                solo = h2d_rate()
        dist.barrier()
        conc = h2d_rate()
        tl = torch.tensor(mine + [solo, conc], device="cuda")
        allv = [torch.zeros_like(tl) for _ in range(world)]
        dist.all_gather(allv, tl)
        link = [{"gpu": i, "pcie_gen": int(v[0].item()), "pcie_width": int(v[1].item()), "h2d_gbs_alone": float(v[2].item()),
                 "h2d_gbs_all_ranks_at_once": float(v[3].item())} for i, v in enumerate(allv)]
        h2d_gbs = conc          # the ceiling of the end-to-end number is the CONCURRENT rate
        tm = torch.tensor([conc], device="cuda"); dist.all_reduce(tm, op=dist.ReduceOp.SUM); h2d_sum = float(tm.item())
    else:
        link = [{"gpu": 0, "pcie_gen": int(mine[0]), "pcie_width": int(mine[1]), "h2d_gbs_alone": h2d_gbs}]
        h2d_sum = h2d_gbs

    # ---- the same link with BOTH directions busy, as in the end-to-end step: the frames going in while the previous results come out
    # (all ranks at once at N > 1).  PCIe is full duplex, the host's memory system is not: when the two directions share it, this is the ceiling.
    bidir_sum = None
    try:
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        cur = torch.cuda.current_stream()

        def bidir_once():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ev0.record()
            s_in.wait_event(ev0); s_out.wait_event(ev0)
            with torch.cuda.stream(s_in):
                d_frames.copy_(h_frames, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_kps.copy_(d_kps, non_blocking=True); h_desc.copy_(d_desc, non_blocking=True)
            cur.wait_stream(s_in); cur.wait_stream(s_out)
            ev1.record(); torch.cuda.synchronize()
            return ev0.elapsed_time(ev1)
        bidir_once()
        t_b = bidir_once()
        tm = torch.tensor([nfr / (t_b * 1e-3)], device="cuda")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.SUM)
        bidir_sum = float(tm.item())
    except Exception:
        bidir_sum = None

    if rank == 0:
        line = {
            "metric": "orb_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": _config(nfr),
            "run": {"chunk_frames_per_launch_set": chunk, "ms_per_step_serial_stage_timed": ms_serial, "host_cpu_affinity_first4": numa,
                    "e2e_steps_timed": e2e_steps},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "call": "eorb_orb_extract_batch (pinned host buffers)",
                    "pcie_h2d_gbs_measured": h2d_gbs, "pcie_bound_frames_per_s": h2d_sum * 1e9 / (W * H),
                    "frac_of_pcie_bound": e2e_value / (h2d_sum * 1e9 / (W * H)),
                    "pcie_bidir_bound_frames_per_s": bidir_sum, "frac_of_pcie_bidir_bound": (e2e_value / bidir_sum) if bidir_sum else None,
                    "host_link": link},
            "gpu_launches": int(launches),
            "roofline": roof,
            "stages": per_stage,
            "cpu_baseline": cpu,
            "cpu_baseline_cv2_assisted": cpu_cv2,
            "parity_checked": {"orb": parity, "hamming": (extra.get("hamming") or {}).get("parity_checked")},
            "strong_scaling": strong,
            "extra": extra,
        }
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _emit(line: dict):
    """the ONE JSON line goes to the real stdout; everything libraries print on fd 1 (e.g. NCCL's version
    banner) has been routed to stderr by _quiet_stdout()"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU per step")
    ap.add_argument("--chunk", type=int, default=CHUNK, help="frames per launch set")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cv2-baseline", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cv2_baseline:
        cores = os.cpu_count() or 1
        _emit(cv2_assisted_baseline(make_batch(min(16 * cores, 1024), 0), cores))
        return 0
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
