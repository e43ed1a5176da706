"""GPU parity: the CUDA ORB extractor (through the C ABI) against the CPU oracle and the golden fixtures.
Integer stages bit-exact; keypoint angles bit-exact (required: <= 1e-3 rad)."""
import ast
import os

import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth

pytestmark = pytest.mark.gpu

ANGLE_TOL_RAD = 1e-3   # BASELINE.json north_star tolerance; the kernels are expected to be bit-exact


def _api():
    from eorb_slam_b200 import api
    return api


def _mk(api, okw, w, h, max_batch=1):
    p = api.ORBxParams(okw["nfeatures"], okw["scale_factor"], okw["nlevels"], okw["ini_th"], okw["min_th"], okw["edge"], (w, h))
    return api.ORBextractor(p, 0, max_batch)


def _orc(okw, w, h):
    return O.OrbOracle(okw["nfeatures"], okw["scale_factor"], okw["nlevels"], okw["ini_th"], okw["min_th"], okw["edge"], w, h)


CFG1 = dict(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, edge=19)


def _compare_stages(ex, orc, nlevels, want_desc=True):
    for l in range(nlevels):
        assert ex.level_size(l) == orc.level_size(l)
        assert np.array_equal(ex.pyramid_level(l), orc.level(l)), "pyramid level %d" % l
    for l in range(nlevels):
        gx, gy, gs = ex.debug_candidates(l)
        ox, oy, os_ = orc.candidates(l)
        assert len(gx) == len(ox), "FAST candidate count level %d: %d vs %d" % (l, len(gx), len(ox))
        # sets compared sorted by (y, x) ...
        go = np.lexsort((gx, gy)); oo = np.lexsort((ox, oy))
        assert np.array_equal(np.stack([gx, gy, gs], 1)[go], np.stack([ox, oy, os_], 1)[oo]), "FAST set level %d" % l
        # ... and the emission order itself equals the reference order (cell-row-major, pixel-row-major)
        assert np.array_equal(gx, ox) and np.array_equal(gy, oy), "FAST order level %d" % l
    for l in range(nlevels):
        gx, gy, gs, ga = ex.debug_level_kps(l)
        ox, oy, os_, oa = orc.level_kps(l)
        assert len(gx) == len(ox), "octree count level %d: %d vs %d" % (l, len(gx), len(ox))
        assert np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gs, os_), "octree selection level %d" % l
        assert np.abs(np.deg2rad(ga - oa)).max(initial=0) <= ANGLE_TOL_RAD
        assert np.array_equal(ga.view(np.uint32), oa.view(np.uint32)), "angle bits level %d" % l
    if want_desc:
        for l in range(nlevels):
            ob = orc.blurred(l)
            if ob is not None:
                assert np.array_equal(ex.debug_blurred(l), ob), "blur level %d" % l


@pytest.mark.parametrize("name", ["cfg1_seed0", "cfg1_seed1_stereo", "cfg1_flat", "mvsec_346x260", "ethz_240x180_e9", "ev_single_level"])
def test_orb_matches_oracle_and_golden(golden_dir, name):
    api = _api()
    g = np.load(os.path.join(golden_dir, "orb_%s.npz" % name))
    fkw = ast.literal_eval(str(g["frame_kw"])); okw = ast.literal_eval(str(g["orb_kw"]))
    img = synth.make_frame(**fkw)
    lap = tuple(int(v) for v in g["lapping"])
    ex = _mk(api, okw, fkw["w"], fkw["h"])
    ret, kps, desc = ex(img, None, lap, True)
    orc = _orc(okw, fkw["w"], fkw["h"])
    oret, okps, odesc = orc.extract(img, lap, True)
    assert list(ex.features_per_level()) == list(orc.features_per_level())
    assert ex.edge_threshold() == orc.edge
    _compare_stages(ex, orc, okw["nlevels"])
    assert ret == oret == int(g["ret"][0])
    assert len(kps) == len(okps)
    assert kps.tobytes() == okps.tobytes() == g["kps"].tobytes(), "keypoints"
    bad = int((desc != odesc).any(axis=1).sum())
    assert bad == 0, "%d descriptor rows differ" % bad
    assert np.array_equal(desc, g["desc"])
    # keypoints-only overload (extDesc = false for event frames, EventFrame.h:38)
    ret2, kps2, d2 = ex(img, None, lap, False)
    assert ret2 == ret and d2 is None and kps2.tobytes() == kps.tobytes()


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_orb_cfg1_three_seeds(seed):
    api = _api()
    img = synth.make_frame(seed)
    ex = _mk(api, CFG1, 752, 480)
    orc = _orc(CFG1, 752, 480)
    ret, kps, desc = ex(img)
    oret, okps, odesc = orc.extract(img)
    _compare_stages(ex, orc, 8)
    assert ret == oret and kps.tobytes() == okps.tobytes() and np.array_equal(desc, odesc)


def test_orb_edge_cases():
    api = _api()
    ex = _mk(api, CFG1, 752, 480)
    ret, kps, desc = ex(np.zeros((0, 0), np.uint8))
    assert ret == -1 and len(kps) == 0                       # return -1 on empty image (ORBextractor.cc:1096)
    ret, kps, desc = ex(np.zeros((480, 752), np.uint8))
    assert ret == 0 and len(kps) == 0 and desc.shape == (0, 32)   # released descriptors
    ret, kps, desc = ex(np.full((480, 752), 255, np.uint8))
    assert len(kps) == 0
    # image size change on the same handle re-plans (the reference sizes its pyramid per call)
    img = synth.make_frame(21, 346, 260)
    ret, kps, desc = ex(img)
    orc = _orc(CFG1, 752, 480)
    oret, okps, odesc = orc.extract(img)
    assert ret == oret and kps.tobytes() == okps.tobytes() and np.array_equal(desc, odesc)
    # non-contiguous rows (stride > width)
    big = np.zeros((480, 800), np.uint8); big[:, :752] = synth.make_frame(22)
    view = big[:, :752]
    img = np.ascontiguousarray(view)
    r1, k1, d1 = ex(img)
    orc2 = _orc(CFG1, 752, 480)
    r2, k2, d2 = orc2.extract(img)
    assert r1 == r2 and k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2)
    # tiny image: upper levels have no FAST grid
    small = synth.make_frame(23, 97, 131)
    exs = _mk(api, dict(CFG1, nfeatures=200), 97, 131)
    orcs = _orc(dict(CFG1, nfeatures=200), 97, 131)
    r1, k1, d1 = exs(small)
    r2, k2, d2 = orcs.extract(small)
    assert r1 == r2 and k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2)


def test_orb_ini_extractor_5x_features():
    # Tracking.cc:369-373 uses 5*nFeatures until initialised: bigger quotas -> bigger octree tables
    api = _api()
    okw = dict(CFG1, nfeatures=5000)
    img = synth.make_frame(31)
    ex = _mk(api, okw, 752, 480)
    orc = _orc(okw, 752, 480)
    ret, kps, desc = ex(img)
    oret, okps, odesc = orc.extract(img)
    _compare_stages(ex, orc, 8, want_desc=False)
    assert ret == oret and kps.tobytes() == okps.tobytes() and np.array_equal(desc, odesc)


def test_orb_batch_matches_single_and_oracle():
    api = _api()
    n = 12
    frames = synth.make_frames(n, seed0=100)
    frames[3] = 0                                   # an empty-result frame inside the batch
    frames[7] = synth.make_frame(7, kind="flat")
    ex = _mk(api, CFG1, 752, 480, max_batch=5)      # 12 frames in chunks of 5,5,2
    kps, desc, nout, mono = ex.extract_batch(frames)
    orc = _orc(CFG1, 752, 480)
    for f in range(n):
        oret, okps, odesc = orc.extract(frames[f])
        assert nout[f] == len(okps) and mono[f] == oret
        assert kps[f, :nout[f]].tobytes() == okps.tobytes(), f
        assert np.array_equal(desc[f, :nout[f]], odesc), f
    tot, counts = O.orb_extract_batch_mt(frames, 4)
    assert list(counts) == list(nout)


def test_orb_batch_device_resident():
    import torch
    api = _api()
    n = 6
    frames = synth.make_frames(n, seed0=200)
    ex = _mk(api, CFG1, 752, 480, max_batch=n)
    cap = ex.cap
    d_img = torch.from_numpy(frames).cuda()
    d_kps = torch.zeros(n * cap * 28, dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros(n * cap * 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(n, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(n, dtype=torch.int32, device="cuda")
    ex.set_stream(torch.cuda.current_stream().cuda_stream)
    ex.extract_batch_raw(d_img.data_ptr(), n, 752, 480, 752, 752 * 480, (0, 1000), True, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                         d_n.data_ptr(), d_mono.data_ptr(), device=True)
    torch.cuda.synchronize()
    nout = d_n.cpu().numpy()
    kps = d_kps.cpu().numpy().view(synth.KEYPOINT_DTYPE).reshape(n, cap)
    desc = d_desc.cpu().numpy().reshape(n, cap, 32)
    orc = _orc(CFG1, 752, 480)
    for f in range(n):
        oret, okps, odesc = orc.extract(frames[f])
        assert nout[f] == len(okps)
        assert kps[f, :nout[f]].tobytes() == okps.tobytes() and np.array_equal(desc[f, :nout[f]], odesc)
    ex.set_stream(None)


@pytest.mark.parametrize("layout", ["unaligned_base", "odd_row_stride", "padded_frames"])
def test_orb_device_input_layouts(layout):
    """device frames that the TMA / word loads cannot read in place (base not 16-byte aligned, row stride not a
    multiple of 16) are staged; padded but aligned layouts are read zero-copy.  Same results either way."""
    import torch
    api = _api()
    n, w, h = 3, 750, 478
    frames = np.stack([synth.make_frame(300 + i, w, h) for i in range(n)])
    if layout == "unaligned_base":
        stride, fstride, off = w, w * h, 3
    elif layout == "odd_row_stride":
        stride, fstride, off = w + 7, (w + 7) * h + 5, 0
    else:
        stride, fstride, off = 768, 768 * 480, 0
    buf = np.zeros(off + n * fstride + 64, np.uint8)
    for i in range(n):
        for y in range(h):
            o = off + i * fstride + y * stride
            buf[o:o + w] = frames[i, y]
    okw = dict(CFG1)
    ex = _mk(api, okw, w, h, max_batch=n)
    cap = ex.cap
    d_buf = torch.from_numpy(buf).cuda()
    d_kps = torch.zeros(n * cap * 28, dtype=torch.uint8, device="cuda"); d_desc = torch.zeros(n * cap * 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(n, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(n, dtype=torch.int32, device="cuda")
    ex.set_stream(torch.cuda.current_stream().cuda_stream)
    ex.extract_batch_raw(d_buf.data_ptr() + off, n, w, h, stride, fstride, (0, 1000), True, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                         d_n.data_ptr(), d_mono.data_ptr(), device=True)
    torch.cuda.synchronize()
    nout = d_n.cpu().numpy()
    kps = d_kps.cpu().numpy().view(synth.KEYPOINT_DTYPE).reshape(n, cap); desc = d_desc.cpu().numpy().reshape(n, cap, 32)
    orc = _orc(okw, w, h)
    for f in range(n):
        _, okps, odesc = orc.extract(frames[f])
        assert nout[f] == len(okps) and kps[f, :nout[f]].tobytes() == okps.tobytes() and np.array_equal(desc[f, :nout[f]], odesc)
    ex.set_stream(None)


def test_tracked_descriptors_and_level_assignment():
    api = _api()
    img = synth.make_frame(41)
    ex = _mk(api, CFG1, 752, 480)
    orc = _orc(CFG1, 752, 480)
    ret, kps, desc = ex(img)
    sub = kps[::7].copy()
    got = ex.ComputeTrackedKPtsDesc(img, sub)
    exp = orc.tracked_desc(img, sub)
    assert np.array_equal(got, exp)
    ref_desc = desc[::7]
    k_gpu = ex.AssignKPtLevelByBestDesc(ref_desc, img, sub)
    k_cpu = orc.assign_level_by_best_desc(ref_desc, img, sub)
    assert np.array_equal(k_gpu["octave"], k_cpu["octave"])


def test_device_math_equals_host_math():
    """shared scalar arithmetic compiled for sm_100a vs the host compile of the same header (toolchain guard)"""
    assert _api().selftest_math(0) == 0


@pytest.mark.parametrize("wh", [(752, 480), (131, 97), (130, 40), (129, 33), (257, 35), (37, 29), (9, 31), (8, 64), (7, 5), (6, 3), (5, 2), (33, 1),
                                (1, 9), (3, 3), (640, 31), (516, 30), (260, 61)])
def test_pyramid_and_blur_odd_sizes(wh):
    """K1/K5 on sizes that exercise partial last words, strips with idle lanes, single-strip levels, bands with a
    tail, rows/columns that bounce more than once (REFLECT_101) — compared with the oracle's cv::resize /
    cv::GaussianBlur restatements level by level (bit-exact)."""
    api = _api()
    w, h = wh
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    okw = dict(CFG1, nlevels=4, nfeatures=100, edge=19)
    ex = _mk(api, okw, w, h)
    ex(img)
    prev = img
    for l in range(4):
        lw, lh = ex.level_size(l)
        if l > 0:
            prev = O.resize_linear(prev, lw, lh)
        got = ex.pyramid_level(l)
        assert got.shape == prev.shape and np.array_equal(got, prev), "pyramid level %d" % l
        assert np.array_equal(ex.debug_blurred(l), O.gauss5(prev)), "blurred level %d" % l


SWEEP = [
    # (w, h, nfeatures, scale, levels, ini, min, edge)  — Appendix A parameter sets and other common sensor shapes
    (346, 260, 1000, 1.1, 16, 10, 1, 30),        # EvMVSEC.yaml
    (240, 180, 1000, 1.07177, 20, 10, 0, 30),    # EvETHZ_EuRoC.yaml
    (346, 260, 1000, 1.26, 6, 10, 1, 15),        # EvMVSEC_ETHZ.yaml (margin 15: descriptor taps reflect)
    (240, 180, 1000, 1.2, 4, 7, 0, 9),           # EvETHZ_SLIDER.yaml
    (640, 480, 1500, 1.2, 8, 20, 7, 19),
    (1241, 376, 2000, 1.2, 8, 20, 7, 19),        # KITTI-shaped: 3 octree roots
    (1280, 720, 1000, 1.2, 8, 20, 7, 19),
    (320, 240, 500, 1.5, 5, 12, 5, 19),
    (853, 481, 1200, 1.33, 7, 25, 25, 21),       # ini == min, odd sizes
    (500, 377, 800, 1.2, 8, 5, 30, 19),          # min > ini (no fallback possible)
    (401, 303, 300, 2.0, 4, 15, 3, 16),
    (752, 480, 100, 1.2, 8, 20, 7, 19),          # tiny quotas
]


@pytest.mark.parametrize("cfg", SWEEP)
def test_orb_parameter_sweep_bit_exact(cfg):
    """geometry sweep: image shapes, pyramid depths / factors, FAST thresholds and margins change every per-level table
    (cell grids, tile alignments, group counts, quotas, octree roots); keypoints and descriptors stay bit-exact."""
    api = _api()
    w, h, nf, sf, nl, ini, mn, edge = cfg
    okw = dict(nfeatures=nf, scale_factor=sf, nlevels=nl, ini_th=ini, min_th=mn, edge=edge)
    img = synth.make_frame(w * 7 + nl, w, h)
    ex = _mk(api, okw, w, h)
    orc = _orc(okw, w, h)
    ret, kps, desc = ex(img)
    oret, okps, odesc = orc.extract(img)
    assert ret == oret and len(kps) == len(okps)
    assert kps.tobytes() == okps.tobytes() and np.array_equal(desc, odesc)
    # a low-texture frame of the same shape exercises the minThFAST fallback cells
    flat = (128 + np.random.default_rng(w + h).integers(-9, 10, (h, w))).astype(np.uint8)
    flat[h // 3:h // 3 + 40, w // 4:w // 4 + 60] += 40
    ret, kps, desc = ex(flat)
    oret, okps, odesc = orc.extract(flat)
    assert ret == oret and kps.tobytes() == okps.tobytes() and np.array_equal(desc, odesc)


def test_extractors_with_different_plans_alternate():
    """EORB-SLAM runs three extractors with different parameters in one process (image ORB, event L1 / L2; Tracking.cc:115-137,
    EvBaseTracker.cpp:163): a later, smaller plan must not undo the kernel attributes an earlier, larger one needs"""
    api = _api()
    img = synth.make_frame(3)
    big = api.ORBextractor(api.ORBxParams(5000, 1.2, 8, 20, 7, 19, (752, 480)))
    r1, k1, d1 = big(img)
    small = api.ORBextractor(api.ORBxParams(400, 1.0, 1, 0, 0, 9, (240, 180)))
    ev = np.ascontiguousarray(img[:180, :240])
    s1 = small(ev, None, (0, 1000), False)
    mid = api.ORBextractor(api.ORBxParams())
    m1 = mid(img)
    for _ in range(2):
        r2, k2, d2 = big(img)
        assert r2 == r1 and k2.tobytes() == k1.tobytes() and np.array_equal(d2, d1)
        s2 = small(ev, None, (0, 1000), False)
        assert s2[0] == s1[0] and s2[1].tobytes() == s1[1].tobytes()
        m2 = mid(img)
        assert m2[1].tobytes() == m1[1].tobytes() and np.array_equal(m2[2], m1[2])
    orc = O.OrbOracle(5000, 1.2, 8, 20, 7, 19, 752, 480)
    _, ok, od = orc.extract(img)
    assert ok.tobytes() == k1.tobytes() and np.array_equal(od, d1)


def test_orb_random_parameter_fuzz():
    """150 random configurations (sizes 64..900 x 48..620, 1..9 levels, scale 1.1..2.0, thresholds, margins incl. the adaptive one,
    1..2500 features, lapping areas, textured / flat frames) against the oracle, bit for bit; documented limits must fail loudly
    (tools/gpu_fuzz_orb.py)"""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_fuzz_orb.py"), "150", "7"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "0 mismatches" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
