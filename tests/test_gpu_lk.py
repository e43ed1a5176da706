"""GPU parity of the pyramidal LK tracker through the C ABI: points, status and error are
BIT-EXACT against the oracle (integer patch sums, identical float operation order) and agree with OpenCV's own
cv2.calcOpticalFlowPyrLK within the freedom its SIMD summation order leaves (see oracle/lk_oracle.cc)."""
import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _api():
    from eorb_slam_b200 import api
    if api.device_count() == 0:
        pytest.fail("GPU tests need a CUDA device; eorb_slam_b200 has no CPU fallback")
    return api


def _pair(seed, w, h, shift=(1.7, -0.9), angle=0.4):
    import cv2
    img = synth.make_frame(seed, w, h)
    M = cv2.getRotationMatrix2D((w / 2, h / 2), angle, 1.0)
    M[0, 2] += shift[0]; M[1, 2] += shift[1]
    return img, cv2.warpAffine(img, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)


def _points(img, n_extra, seed):
    import cv2
    h, w = img.shape
    rng = np.random.default_rng(seed)
    corners = cv2.goodFeaturesToTrack(img, 400, 0.01, 5).reshape(-1, 2)
    extra = np.stack([rng.uniform(-8, w + 8, n_extra), rng.uniform(-8, h + 8, n_extra)], 1)
    return np.concatenate([corners, extra]).astype(np.float32)


@pytest.mark.parametrize("cfg", [dict(w=240, h=180, win=23, lv=1, it=10, eps=0.03),      # EvETHZ.yaml:205-208
                                 dict(w=346, h=260, win=23, lv=1, it=10, eps=0.03),      # MVSEC-shaped frames
                                 dict(w=240, h=180, win=15, lv=3, it=30, eps=0.01),
                                 dict(w=752, h=480, win=21, lv=2, it=10, eps=0.03),
                                 dict(w=97, h=61, win=23, lv=4, it=10, eps=0.03)])       # pyramid stops early (level <= window)
def test_lk_bit_exact_vs_oracle_and_close_to_cv2(cfg):
    import cv2
    api = _api()
    w, h, win, lv = cfg["w"], cfg["h"], cfg["win"], cfg["lv"]
    img, nxt = _pair(w + win, w, h)
    pts = _points(img, 80, w * 5 + win)
    tr = api.ELK_Tracker(win, lv, cfg["it"], cfg["eps"], max_size=(w, h), max_points=len(pts))
    assert tr.setRefImage(img, pts) == 0
    got_p, got_s, got_e = tr.trackCurrImage(nxt)
    exp_p, exp_s, exp_e, exp_lv = O.lk_track(img, nxt, pts, None, win, lv, cfg["it"], cfg["eps"])
    assert tr.levels_used == exp_lv
    assert np.array_equal(got_s, exp_s)
    assert got_p.tobytes() == exp_p.tobytes(), float(np.abs(got_p - exp_p).max())
    assert got_e.tobytes() == exp_e.tobytes()
    # and against OpenCV itself
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, cfg["it"], cfg["eps"])
    ref_p, ref_s, _ = cv2.calcOpticalFlowPyrLK(img, nxt, pts.reshape(-1, 1, 2), None, winSize=(win, win), maxLevel=lv, criteria=crit)
    ref_p = ref_p.reshape(-1, 2); ref_s = ref_s.reshape(-1)
    same = got_s == ref_s
    assert same.mean() >= 0.99
    ok = same & (ref_s == 1)
    d = np.abs(got_p[ok] - ref_p[ok]).max(axis=1)
    assert (d <= 0.02).mean() >= 0.99 and ok.sum() > 50       # tolerance: 0.02 px on >= 99 % of the tracked points

    # OPTFLOW_USE_INITIAL_FLOW with the previously tracked points, second frame (KLT_Tracker.cpp:63-65)
    _, nxt2 = _pair(w + win, w, h, shift=(3.1, -1.6), angle=0.8)
    got_p2, got_s2, got_e2 = tr.trackCurrImage(nxt2, got_p)
    exp_p2, exp_s2, exp_e2, _ = O.lk_track(img, nxt2, pts, got_p, win, lv, cfg["it"], cfg["eps"])
    assert np.array_equal(got_s2, exp_s2) and got_p2.tobytes() == exp_p2.tobytes() and got_e2.tobytes() == exp_e2.tobytes()


def test_lk_on_event_frames_and_edge_cases():
    """the reference's use: keypoints of an event frame tracked into the next event frame (EvAsynchTracker.cpp:590-594)"""
    api = _api()
    per, w, h = 2000, 240, 180
    ev = synth.make_events(per * 2, seed=7, w=w, h=h)
    f0, _, _ = O.ev_accumulate(ev[:per], w, h, 1.0, mode=1)
    f1, _, _ = O.ev_accumulate(ev[per // 2:per + per // 2], w, h, 1.0, mode=1)
    i0, i1 = O.normalize_minmax_u8(f0), O.normalize_minmax_u8(f1)
    orc = O.OrbOracle(400, 1.0, 1, 0, 0, 9, w, h)
    _, kps, _ = orc.extract(i0, (0, 1000), False)
    assert len(kps) > 50
    tr = api.ELK_Tracker(max_size=(w, h), max_points=len(kps))
    assert tr.setRefImage(i0, kps) == 0
    p, s, e = tr.trackCurrImage(i1)
    pts = np.stack([kps["x"], kps["y"]], 1)
    ep, es, ee, _ = O.lk_track(i0, i1, pts)
    assert np.array_equal(s, es) and p.tobytes() == ep.tobytes() and e.tobytes() == ee.tobytes()
    # tracking a frame onto itself: every trackable point stays put
    p0, s0, _ = tr.trackCurrImage(i0)
    assert np.abs(p0[s0 == 1] - pts[s0 == 1]).max() < 1e-3
    # empty inputs: EORB_EMPTY (-1), like the assert at KLT_Tracker.cpp:24; tracking before set_ref is a state error
    tr2 = api.ELK_Tracker(max_size=(w, h), max_points=16)
    assert tr2.setRefImage(i0, np.zeros((0, 2), np.float32)) == -1
    with pytest.raises(api.EorbError):
        tr2.n = 1; tr2.shape = i0.shape
        tr2.trackCurrImage(i0)


@pytest.mark.parametrize("init", [False, True])
def test_track_and_match_state_stays_on_device(init):
    """trackAndMatchCurrImage / trackAndMatchCurrImageInit (KLT_Tracker.cpp:215-242) over three event-sized frames: the tracked points
    stay in HBM as the next call's initial flow, refineTrackedPts (+ refineFirstOctaveLevel) runs on the device; every output of every
    frame equals the oracle chain (LK oracle -> refine oracle) bit for bit, with the caller's vectors carried from frame to frame."""
    api = _api()
    w, h = 240, 180
    img, _ = _pair(5, w, h)
    frames = [_pair(5, w, h, shift=(0.8 * k, -0.5 * k), angle=0.2 * k)[1] for k in (1, 2, 3)]
    pts = _points(img, 60, 77)
    rng = np.random.default_rng(3)
    ref = np.zeros(len(pts), api.KEYPOINT_DTYPE)
    ref["x"] = pts[:, 0]; ref["y"] = pts[:, 1]; ref["size"] = 31; ref["angle"] = rng.uniform(0, 360, len(pts))
    ref["response"] = rng.integers(1, 100, len(pts)); ref["octave"] = rng.integers(0, 3, len(pts)); ref["class_id"] = -1
    tr = api.ELK_Tracker(23, 1, 10, 0.03, max_size=(w, h), max_points=len(pts) + 5)
    assert tr.setRefImageKPts(img, ref) == 0
    last = pts.copy()
    m12 = cnt = em12 = ecnt = None
    for k, f in enumerate(frames):
        nm, tk, m12, cnt, disp = tr.trackAndMatchCurrImage(f, m12, cnt, init=init)
        ep, es, _, _ = O.lk_track(img, f, pts, last, 23, 1, 10, 0.03)
        enm, etk, em12, ecnt, edisp = O.lk_refine(ep, es, ref, w, h, init, em12, ecnt)
        assert nm == enm and tk.tobytes() == etk.tobytes(), k
        assert np.array_equal(m12, em12) and np.array_equal(cnt, ecnt) and disp.tobytes() == edisp.tobytes(), k
        assert tr.getLastTrackedPts().tobytes() == ep.tobytes()
        last = ep
    assert 0 < nm < len(pts)
    # setLastTrackedPts with a list of another size: the next call starts from the reference points without initial flow (:63-70)
    tr.setLastTrackedPts(ref[:10])
    nm, tk, _, _, _ = tr.trackAndMatchCurrImage(frames[0], init=init)
    ep, es, _, _ = O.lk_track(img, frames[0], pts, None, 23, 1, 10, 0.03)
    enm, etk, _, _, _ = O.lk_refine(ep, es, ref, w, h, init)
    assert nm == enm and tk.tobytes() == etk.tobytes()
    # ... and with a full list it is the initial flow
    moved = ref.copy(); moved["x"] += 1.0
    tr.setLastTrackedPts(moved)
    nm, tk, _, _, _ = tr.trackAndMatchCurrImage(frames[1], init=init)
    ep, es, _, _ = O.lk_track(img, frames[1], pts, np.stack([moved["x"], moved["y"]], 1), 23, 1, 10, 0.03)
    enm, etk, _, _, _ = O.lk_refine(ep, es, ref, w, h, init)
    assert nm == enm and tk.tobytes() == etk.tobytes()


def test_track_and_match_without_reference_is_empty():
    api = _api()
    tr = api.ELK_Tracker(max_size=(64, 64), max_points=8)
    tr.n = 0; tr.shape = (64, 64)
    assert tr.trackAndMatchCurrImage(np.zeros((64, 64), np.uint8))[0] == 0


def test_track_and_match_device_variant_equals_host_variant():
    """eorb_lk_track_and_match_device: frame and outputs resident on the device (the event front end's frames are produced there) ==
    the host-buffer call, byte for byte, over two frames"""
    import torch
    api = _api()
    w, h = 240, 180
    img, _ = _pair(9, w, h)
    frames = [_pair(9, w, h, shift=(0.7 * k, 0.4 * k), angle=0.15 * k)[1] for k in (1, 2)]
    pts = _points(img, 40, 5)
    ref = np.zeros(len(pts), api.KEYPOINT_DTYPE)
    ref["x"] = pts[:, 0]; ref["y"] = pts[:, 1]; ref["size"] = 31; ref["angle"] = 10.0; ref["response"] = 50; ref["octave"] = np.arange(len(pts)) % 2
    ref["class_id"] = -1
    n = len(pts)
    a = api.ELK_Tracker(23, 1, 10, 0.03, max_size=(w, h), max_points=n)
    b = api.ELK_Tracker(23, 1, 10, 0.03, max_size=(w, h), max_points=n)
    assert a.setRefImageKPts(img, ref) == 0 and b.setRefImageKPts(img, ref) == 0
    st = torch.cuda.Stream()
    b.set_stream(st.cuda_stream)
    d_tr = torch.zeros(n * 28, dtype=torch.uint8, device="cuda"); d_m = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_disp = torch.zeros(n, dtype=torch.float32, device="cuda"); d_c = torch.zeros(2, dtype=torch.int32, device="cuda")
    for f in frames:
        nm, tk, m12, cnt, disp = a.trackAndMatchCurrImage(f, init=True)
        d_f = torch.from_numpy(f).cuda()
        torch.cuda.synchronize()
        b.trackAndMatchCurrImage_device(d_f.data_ptr(), w, d_tr.data_ptr(), d_m.data_ptr(), d_disp.data_ptr(), d_c.data_ptr(), init=True)
        st.synchronize()
        c = d_c.cpu().numpy()
        assert int(c[0]) == nm and int(c[1]) == len(disp)
        assert d_tr.cpu().numpy().tobytes() == tk.tobytes()
        assert d_disp.cpu().numpy()[:c[1]].tobytes() == disp.tobytes()
        kept = (d_m.cpu().numpy() & 2) != 0
        assert np.array_equal(kept, m12 == np.arange(n))
    b.set_stream(None)
