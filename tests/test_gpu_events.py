"""GPU parity: event frames against the CPU oracle.  Float sums are order-dependent (atomics), so the bar is
max |gpu - oracle| <= 1e-4 * peak (BASELINE.json north_star); the u8 frame may differ by 1 LSB."""
import os

import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4
K_ETHZ = (199.09, 198.83, 132.19, 110.71)
K_MVSEC = (226.38, 226.15, 173.65, 133.73)


def _api():
    from eorb_slam_b200 import api
    return api


def _close(gpu, ref):
    peak = float(np.abs(ref).max())
    err = float(np.abs(gpu - ref).max())
    assert err <= REL_TOL * max(peak, 1e-12), (err, peak)


def test_gauss_frames_match_oracle_and_golden(golden_dir):
    api = _api()
    g = np.load(os.path.join(golden_dir, "events_2000.npz"))
    ev = synth.make_events(2000, seed=int(g["seed"][0]), w=240, h=180)
    cv = api.EvImConverter(0, 1, 100000, 346, 260)
    img, u8 = cv.ev2im_gauss(ev, 240, 180, 1.0, False, True, both=True)
    _close(img, g["gauss"])
    assert np.abs(u8.astype(int) - g["gauss_u8"].astype(int)).max() <= 1
    assert (u8 != g["gauss_u8"]).mean() < 0.01
    imgn = cv.ev2im(ev, 240, 180, False, False)
    _close(imgn, g["nearest"])
    img3 = cv.ev2mci_gg_f(ev, tuple(g["K"]), g["Trot"], 1.0, 240, 180, 1.0, False, False)
    _close(img3, g["se3"])
    img4 = cv.ev2mci_gg_f_2d(ev, tuple(g["K"]), g["se2_params"], 240, 180, 1.0, False, False)
    _close(img4, g["se2"])


@pytest.mark.parametrize("n,w,h,sigma", [(2000, 240, 180, 1.0), (6000, 240, 180, 1.0), (50000, 346, 260, 1.0), (3000, 240, 180, 2.0),
                                         (1, 240, 180, 1.0), (500, 64, 48, 0.5)])
def test_gauss_sizes(n, w, h, sigma):
    api = _api()
    ev = synth.make_events(n, seed=n + w, w=w, h=h)
    cv = api.EvImConverter(0, 1, 100000, 346, 260)
    img = cv.ev2im_gauss(ev, w, h, sigma, False, False)
    ref, (mn, mx), _ = O.ev_accumulate(ev, w, h, sigma, mode=1)
    _close(img, ref)
    # polarity variant
    imgp = cv.ev2im_gauss(ev, w, h, sigma, True, False)
    refp, _, _ = O.ev_accumulate(ev, w, h, sigma, mode=1, pol=True)
    _close(imgp, refp)


def test_motion_compensation_mvsec_shape():
    api = _api()
    ev = synth.make_events(50000, seed=77, w=346, h=260, mean_dt=2e-7)
    cv = api.EvImConverter(0, 1, 100000, 346, 260)
    for omega, t in (([0.01, -0.02, 0.03], (0, 0, 0)), ([0.0, 0.0, 0.0], (0, 0, 0)), ([0.02, 0.01, -0.015], (0.01, -0.02, 0.005)),
                     ([2.5, 0.3, -1.0], (0, 0, 0))):   # last: large rotation -> trace <= 0 branch of the quaternion conversion
        T = synth.rotation_tcw(omega, t)
        img = cv.ev2mci_gg_f(ev, K_MVSEC, T, 1.0, 346, 260, 1.0, False, False)
        ref, _, _ = O.ev_accumulate(ev, 346, 260, 1.0, mode=2, Tcw=T, depth=1.0, K=np.array(K_MVSEC, np.float32))
        _close(img, ref)
    # live callers: normalized=false then cv::normalize(NORM_MINMAX) (EvImBuilder.cpp:969-976)
    T = synth.rotation_tcw([0.01, -0.02, 0.03])
    img, u8 = cv.ev2mci_gg_f(ev, K_MVSEC, T, 1.0, 346, 260, 1.0, False, False, both=True, norm_mode=api.NORM_MINMAX)
    ref, _, _ = O.ev_accumulate(ev, 346, 260, 1.0, mode=2, Tcw=T, depth=1.0, K=np.array(K_MVSEC, np.float32))
    ref8 = O.normalize_minmax_u8(ref)
    assert np.abs(u8.astype(int) - ref8.astype(int)).max() <= 1
    # SE2 with scale
    for se2 in ([0.03, 0.01, -0.02], [0.02, -0.01, 0.015, 0.97]):
        img = cv.ev2mci_gg_f_2d(ev, K_MVSEC, se2, 346, 260, 1.0, False, False)
        ref, _, _ = O.ev_accumulate(ev, 346, 260, 1.0, mode=3, K=np.array(K_MVSEC, np.float32), se2=np.array(se2, np.float32))
        _close(img, ref)


def test_motion_compensation_kannala_brandt8_camera():
    """the camera MVSEC actually uses (Examples/Event/EvMVSEC.yaml:50-63): KannalaBrandt8 unproject / project inside the SE3 and SE2
    warps, against the oracle (which equals the reference's own KannalaBrandt8.cpp bit for bit, tests/test_ref_pin.py)"""
    api = _api()
    K8 = (226.38018519795807, 226.15002947047415, 173.6470807871759, 133.73271487507847,
          -0.048031442223833355, 0.011330957517194437, -0.055378166304281135, 0.021500973881459395)
    ev = synth.make_events(50000, seed=78, w=346, h=260, mean_dt=2e-7)
    cv = api.EvImConverter(0, 1, 100000, 346, 260)
    for omega, t in (([0.01, -0.02, 0.03], (0, 0, 0)), ([0.02, 0.01, -0.015], (0.01, -0.02, 0.005))):
        T = synth.rotation_tcw(omega, t)
        img = cv.ev2mci_gg_f(ev, K8, T, 1.2, 346, 260, 1.0, False, False)
        ref, _, _ = O.ev_accumulate(ev, 346, 260, 1.0, mode=2, Tcw=T, depth=1.2, K=K8[:4], kb8=K8[4:])
        _close(img, ref)
        pin, _, _ = O.ev_accumulate(ev, 346, 260, 1.0, mode=2, Tcw=T, depth=1.2, K=K8[:4])
        assert np.abs(ref - pin).max() > 0.5
    for se2 in ([0.03, 0.01, -0.02], [0.02, -0.01, 0.015, 0.97]):
        img = cv.ev2mci_gg_f_2d(ev, K8, se2, 346, 260, 1.0, False, False)
        ref, _, _ = O.ev_accumulate(ev, 346, 260, 1.0, mode=3, K=K8[:4], kb8=K8[4:], se2=np.array(se2, np.float32))
        _close(img, ref)


def test_empty_and_out_of_image():
    api = _api()
    cv = api.EvImConverter(0, 1, 1000, 240, 180)
    ev0 = np.zeros(0, synth.EVENT_DTYPE)
    img = cv.ev2im_gauss(ev0, 240, 180, 1.0, False, False)
    assert img.shape == (180, 240) and not img.any()
    out = np.zeros(5, synth.EVENT_DTYPE)
    out["x"] = [-50, 500, 10, -1.2, 239.9]; out["y"] = [10, 10, -70, -1.5, 179.9]; out["ts"] = np.arange(5) * 1e-6
    img = cv.ev2im_gauss(out, 240, 180, 1.0, False, False)
    ref, _, _ = O.ev_accumulate(out, 240, 180, 1.0, mode=1)
    _close(img, ref)
    u8 = cv.ev2im_gauss(out[:3], 240, 180, 1.0, False, True)     # nothing lands in the image -> zeros
    assert not u8.any()


def test_windows_batch_device_then_orb_on_event_frames():
    """config 2: fixed-size windows (EvTrackManager.cpp:272-286) -> event frames -> ORB on the frames.
    ORB parity is checked on the ORACLE's u8 frame fed to both paths (the u8 frame may differ by 1 LSB)."""
    import torch
    api = _api()
    nwin, per = 8, 2000
    ev = synth.make_events(nwin * per, seed=5, w=240, h=180)
    cv = api.EvImConverter(0, nwin, nwin * per, 240, 180)
    d_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1)).cuda()
    d_img = torch.zeros(nwin * 180 * 240, dtype=torch.float32, device="cuda")
    d_u8 = torch.zeros(nwin * 180 * 240, dtype=torch.uint8, device="cuda")
    cv.set_stream(torch.cuda.current_stream().cuda_stream)
    p = cv.make_params(api.EV_GAUSS, 240, 180, 1.0, False, api.NORM_RUNNING)
    offs = np.arange(nwin + 1, dtype=np.int64) * per
    cv.accumulate_batch_device(d_ev.data_ptr(), offs, p, d_img.data_ptr(), d_u8.data_ptr())
    torch.cuda.synchronize()
    imgs = d_img.cpu().numpy().reshape(nwin, 180, 240); u8s = d_u8.cpu().numpy().reshape(nwin, 180, 240)
    # L1 event extractor: single level, N=400, FAST 0/0, margin 9, keypoints only (EvETHZ.yaml:185-199)
    ex = api.ORBextractor(api.ORBxParams(400, 1.0, 1, 0, 0, 9, (240, 180)))
    orc = O.OrbOracle(400, 1.0, 1, 0, 0, 9, 240, 180)
    for i in range(nwin):
        ref, (mn, mx), ref8 = O.ev_accumulate(ev[i * per:(i + 1) * per], 240, 180, 1.0, mode=1, normalize=True)
        _close(imgs[i], ref)
        assert np.abs(u8s[i].astype(int) - ref8.astype(int)).max() <= 1
        r1, k1, _ = ex(ref8, None, (0, 1000), False)
        r2, k2, _ = orc.extract(ref8, (0, 1000), False)
        assert r1 == r2 and k1.tobytes() == k2.tobytes()
    cv.set_stream(None)


def test_windows_batch_host_entry_matches_oracle_and_device_entry():
    """eorb_ev_accumulate_batch (host events + offsets in, frames back on the host): ragged windows incl. an empty one, float and u8 outputs,
    per-window SE3 poses; == the oracle per window (tolerance of the file) and == the device entry point bit for bit"""
    import torch
    api = _api()
    sizes = [2000, 0, 1500, 3100, 1, 2400]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    ev = synth.make_events(int(offs[-1]), seed=15, w=240, h=180)
    nwin = len(sizes)
    cv = api.EvImConverter(0, nwin, int(offs[-1]), 240, 180)
    p = cv.make_params(api.EV_GAUSS, 240, 180, 1.0, False, api.NORM_RUNNING)
    f32 = np.zeros((nwin, 180, 240), np.float32); u8 = np.zeros((nwin, 180, 240), np.uint8)
    cv.accumulate_batch(ev, offs, p, f32, u8)
    d_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1)).cuda()
    d_img = torch.zeros(nwin * 180 * 240, dtype=torch.float32, device="cuda"); d_u8 = torch.zeros(nwin * 180 * 240, dtype=torch.uint8, device="cuda")
    cv.accumulate_batch_device(d_ev.data_ptr(), offs, p, d_img.data_ptr(), d_u8.data_ptr())
    cv.synchronize()
    assert f32.tobytes() == d_img.cpu().numpy().tobytes() and u8.tobytes() == d_u8.cpu().numpy().tobytes()
    for i, n in enumerate(sizes):
        if n == 0:
            assert not f32[i].any() and not u8[i].any()
            continue
        ref, _, ref8 = O.ev_accumulate(ev[offs[i]:offs[i + 1]], 240, 180, 1.0, mode=1, normalize=True)
        _close(f32[i], ref)
        assert np.abs(u8[i].astype(int) - ref8.astype(int)).max() <= 1
    # float frames only, motion-compensated with one pose per window
    K = (199.09, 198.83, 132.19, 110.71)
    poses = np.tile(np.eye(4, dtype=np.float32), (nwin, 1, 1))
    for i in range(nwin):
        a = 0.01 * (i + 1)
        poses[i, :3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
    pm = cv.make_params(api.EV_SE3, 240, 180, 1.0, False, api.NORM_NONE, Tcw=poses[0], medDepth=1.0, camera=K)
    g32 = np.zeros((nwin, 180, 240), np.float32)
    cv.accumulate_batch(ev, offs, pm, g32, None, poses=poses.reshape(nwin, 16))
    for i, n in enumerate(sizes):
        if n < 2:
            continue
        ref, _, _ = O.ev_accumulate(ev[offs[i]:offs[i + 1]], 240, 180, 1.0, mode=2, Tcw=poses[i], depth=1.0, K=np.array(K, np.float32))
        _close(g32[i], ref)
    with pytest.raises(Exception):
        cv.accumulate_batch(ev, offs, pm, None, u8)          # a u8 output needs a normalisation mode
    # a large packet takes the pipelined path (four window ranges on two streams): == the device entry point bit for bit, oracle on samples;
    # plain, motion-compensated (per-window poses) and the order-dependent pol = true normalisation
    rng = np.random.default_rng(3)
    sizes = [int(v) for v in rng.integers(2500, 4500, 26)]
    sizes[5] = 0; sizes[17] = 1
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    nwin = len(sizes)
    assert offs[-1] >= 65536
    ev = synth.make_events(int(offs[-1]), seed=16, w=240, h=180)
    cv2_ = api.EvImConverter(0, nwin, int(offs[-1]), 240, 180)
    d_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1)).cuda()
    poses = np.tile(np.eye(4, dtype=np.float32), (nwin, 1, 1))
    for i in range(nwin):
        a = 0.002 * (i + 1)
        poses[i, :3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
    variants = [cv2_.make_params(api.EV_GAUSS, 240, 180, 1.0, False, api.NORM_RUNNING),
                cv2_.make_params(api.EV_SE3, 240, 180, 1.0, False, api.NORM_MINMAX, Tcw=poses[0], medDepth=1.0, camera=K),
                cv2_.make_params(api.EV_GAUSS, 240, 180, 1.0, True, api.NORM_RUNNING)]
    for vi, pv in enumerate(variants):
        f32 = np.zeros((nwin, 180, 240), np.float32); u8 = np.zeros((nwin, 180, 240), np.uint8)
        ps = poses.reshape(nwin, 16) if vi == 1 else None
        cv2_.accumulate_batch(ev, offs, pv, f32, u8, poses=ps)
        d_img = torch.zeros(nwin * 180 * 240, dtype=torch.float32, device="cuda"); d_u8 = torch.zeros(nwin * 180 * 240, dtype=torch.uint8, device="cuda")
        cv2_.accumulate_batch_device(d_ev.data_ptr(), offs, pv, d_img.data_ptr(), d_u8.data_ptr(), poses=ps)
        cv2_.synchronize()
        assert f32.tobytes() == d_img.cpu().numpy().tobytes() and u8.tobytes() == d_u8.cpu().numpy().tobytes(), vi
        if vi == 0:
            for i in (0, 6, 13, 25):
                ref, _, ref8 = O.ev_accumulate(ev[offs[i]:offs[i + 1]], 240, 180, 1.0, mode=1, normalize=True)
                _close(f32[i], ref)
                assert np.abs(u8[i].astype(int) - ref8.astype(int)).max() <= 1
            assert not f32[5].any()


def test_config5_mvsec_mc_frames_orb_and_frame_to_frame_matching():
    """configs[4]: MVSEC-shaped 346x260, 50k events/window, motion-compensated frames from a synthetic rotation
    (t = 0, medDepth = 1) -> cv::normalize(MINMAX) -> ORB (EvMVSEC_ETHZ.yaml values, margin raised to 19 for the
    descriptor-parity run) -> brute-force best-2 + ratio 0.9 + TH_LOW + rotation histogram between consecutive
    windows.  Event frames are toleranced; everything downstream is bit-exact on the oracle's u8 frames."""
    api = _api()
    per, nwin = 50000, 3
    ev = synth.make_events(per * nwin, seed=91, w=346, h=260, mean_dt=2e-7, n_edges=60)
    cv = api.EvImConverter(0, 1, per, 346, 260)
    ex = api.ORBextractor(api.ORBxParams(1000, 1.26, 6, 10, 1, 19, (346, 260)))
    orc = O.OrbOracle(1000, 1.26, 6, 10, 1, 19, 346, 260)
    Kc = np.array(K_MVSEC, np.float32)
    frames = []
    for i in range(nwin):
        w_ev = ev[i * per:(i + 1) * per]
        dt = float(w_ev["ts"][-1] - w_ev["ts"][0])
        omega = np.array([0.8, -1.1, 2.0]) * dt * (1 + 0.2 * i)          # |omega| <= 3 rad/s over the window's DT
        T = synth.rotation_tcw(omega)
        img, u8 = cv.ev2mci_gg_f(w_ev, K_MVSEC, T, 1.0, 346, 260, 1.0, False, False, both=True, norm_mode=api.NORM_MINMAX)
        ref, _, _ = O.ev_accumulate(w_ev, 346, 260, 1.0, mode=2, Tcw=T, depth=1.0, K=Kc)
        _close(img, ref)
        ref8 = O.normalize_minmax_u8(ref)
        assert np.abs(u8.astype(int) - ref8.astype(int)).max() <= 1
        r1, k1, d1 = ex(ref8)
        r2, k2, d2 = orc.extract(ref8)
        assert r1 == r2 and k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2)
        assert len(k1) > 100
        frames.append((k1, d1))
    m = api.ORBmatcher(0.9, True)
    for (ka, da), (kb, db) in zip(frames[:-1], frames[1:]):
        n, m12 = m.SearchBruteForce(da, db, ka["angle"], kb["angle"])
        exp = O.hamming_best2(da, db, 50, 0.9)
        e12 = np.where(exp["accepted"] == 1, exp["best_idx"], -1).astype(np.int32)
        n_ref, m_ref = O.rotation_filter(ka["angle"], kb["angle"], e12)
        assert n == n_ref and np.array_equal(m12, m_ref)


def test_contrast_metric_and_best_candidate_selection_on_device():
    """measureImageFocus* / imageMeanLocal of event frames (EventConversion.cc:79-162) and the
    best-of-N motion-compensated candidate choice (EvImBuilder.cpp:1206-1215) without the frames leaving the GPU.
    Tolerance 2e-6 relative (double sums in a different order, cast to float once per cell)."""
    import torch
    api = _api()
    w, h, per = 346, 260, 20000
    ev = synth.make_events(per, seed=17, w=w, h=h, mean_dt=2e-7, n_edges=40)
    cv = api.EvImConverter(0, 4, per * 4, w, h)
    ref, _, _ = O.ev_accumulate(ev, w, h, 1.0, mode=1)
    for what in (0, 1, 2):
        for avg in (True, False):
            got = cv._focus(ref, what, avg)
            exp = O.image_focus(ref, what, avg)
            assert abs(got - exp) <= 2e-6 * max(abs(exp), 1e-3), (what, avg, got, exp)
    assert cv.measureImageFocus(ref) == cv.measureImageFocusLocal(ref, True)
    small = (np.random.default_rng(3).random((47, 61)) * 3).astype(np.float32)          # partial cells on both edges
    assert abs(cv.measureImageFocus(small) - O.image_focus(small)) <= 2e-6 * O.image_focus(small)
    # four motion-compensation candidates (different rotation hypotheses) of the same window, frames stay on the device
    dt = float(ev["ts"][-1] - ev["ts"][0])
    hyps = [np.array([0.0, 0.0, 0.0]), np.array([0.5, -0.7, 1.5]) * dt, np.array([-1.0, 0.4, -2.0]) * dt, np.array([0.2, 0.1, 3.0]) * dt]
    poses = np.stack([synth.rotation_tcw(o) for o in hyps]).astype(np.float32).reshape(4, 16)
    d_ev = torch.from_numpy(np.tile(ev.view(np.uint8).reshape(-1), 4)).cuda()
    d_img = torch.empty(4 * h * w, dtype=torch.float32, device="cuda")
    p = cv.make_params(api.EV_SE3, w, h, 1.0, False, api.NORM_NONE, medDepth=1.0, camera=K_MVSEC)
    offs = np.arange(5, dtype=np.int64) * per
    cv.accumulate_batch_device(d_ev.data_ptr(), offs, p, d_img.data_ptr(), None, poses=poses)
    focus = cv.image_focus_device(d_img.data_ptr(), 4, w, h)
    frames = d_img.cpu().numpy().reshape(4, h, w)
    exp = np.array([O.image_focus(f) for f in frames], np.float32)
    assert np.all(np.abs(focus - exp) <= 2e-6 * exp)
    assert int(np.argmax(focus)) == int(np.argmax(exp))


@pytest.mark.parametrize("cfg", [dict(w=240, h=180, n=6000, K=(199.09, 198.83, 132.19, 110.71)), dict(w=346, h=260, n=30000, K=K_MVSEC)])
def test_mci_jacobian_matches_oracle(cfg):
    """SURVEY §8f rank 2 (second half): ev2mci_gg_f_jac (EventConversion.cc:533-662).  Seven splat images + six product
    means; float sums in another order, hence toleranced: 2e-4 of the largest component (measured ~1e-6)."""
    api = _api()
    w, h, n, K = cfg["w"], cfg["h"], cfg["n"], cfg["K"]
    ev = synth.make_events(n, seed=n, w=w, h=h, mean_dt=2e-7, n_edges=40)
    dt = float(ev["ts"][-1] - ev["ts"][0])
    T = synth.rotation_tcw(np.array([0.5, -0.7, 1.5]) * dt).astype(np.float64)
    R, t = T[:3, :3], np.array([0.02, -0.01, 0.03])
    cv = api.EvImConverter(0, 1, n, w, h)
    for glob in (False, True):
        for pol in (False, True):
            got = cv.ev2mci_gg_f_jac(ev, K, R, t, 1.3, w, h, 1.0, pol, glob)
            exp = O.ev_mci_jac(ev, w, h, 1.0, R, t, 1.3, K, pol, glob)
            assert np.abs(got - exp).max() <= 2e-4 * np.abs(exp).max(), (glob, pol, got, exp)
    # no events: zero Jacobian and EORB_EMPTY, like the reference's early return (:543-546)
    assert np.all(cv.ev2mci_gg_f_jac(ev[:0], K, R, t, 1.0, w, h, 1.0) == 0)


def test_config2_chain_stays_on_the_device():
    """configs[1] end to end without a host round trip: event windows -> Gaussian event frames (u8, in HBM) -> single-level event
    extractor on those frames (eorb_orb_extract_batch_device).  The extractor's output must equal the oracle's extraction of the
    very frames the device produced (downloaded only for the check)."""
    import torch
    api = _api()
    nwin, per, w, h = 16, 2000, 240, 180
    ev = synth.make_events(nwin * per, seed=8, w=w, h=h)
    cv = api.EvImConverter(0, nwin, nwin * per, w, h)
    st = torch.cuda.current_stream().cuda_stream
    cv.set_stream(st)
    d_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1)).cuda()
    d_img = torch.zeros(nwin * h * w, dtype=torch.float32, device="cuda"); d_u8 = torch.zeros(nwin * h * w, dtype=torch.uint8, device="cuda")
    p = cv.make_params(api.EV_GAUSS, w, h, 1.0, False, api.NORM_RUNNING)
    cv.accumulate_batch_device(d_ev.data_ptr(), np.arange(nwin + 1, dtype=np.int64) * per, p, d_img.data_ptr(), d_u8.data_ptr())
    ex = api.ORBextractor(api.ORBxParams(400, 1.0, 1, 0, 0, 9, (w, h)), 0, nwin)
    ex.set_stream(st)
    cap = ex.cap
    d_kps = torch.zeros(nwin * cap * 28, dtype=torch.uint8, device="cuda"); d_desc = torch.zeros(nwin * cap * 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(nwin, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(nwin, dtype=torch.int32, device="cuda")
    ex.extract_batch_raw(d_u8.data_ptr(), nwin, w, h, w, w * h, (0, 1000), False, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr(),
                         d_mono.data_ptr(), device=True)
    torch.cuda.synchronize()
    frames = d_u8.cpu().numpy().reshape(nwin, h, w)
    ns = d_n.cpu().numpy(); kps = d_kps.cpu().numpy().view(synth.KEYPOINT_DTYPE).reshape(nwin, cap)
    orc = O.OrbOracle(400, 1.0, 1, 0, 0, 9, w, h)
    for i in range(nwin):
        _, ok, _ = orc.extract(frames[i], (0, 1000), False)
        assert ns[i] == len(ok) and kps[i][:ns[i]].tobytes() == ok.tobytes(), i
    assert ns.min() > 50
    cv.set_stream(None); ex.set_stream(None)


def test_random_parameter_fuzz_of_the_other_rows():
    """random sizes / parameters for the event-frame overloads (incl. sigma != 1, single-event windows), the LK tracker, both
    guided matchers and the vocabulary transform against the oracle (tools/gpu_fuzz_rest.py)"""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_fuzz_rest.py"), "40", "11"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "0 mismatches" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("mode", ["nearest", "gauss", "se3", "se2"])
def test_polarity_frames_with_running_normalisation_follow_the_event_order(mode):
    """pol = true + normalized = true: the reference normalises with the RUNNING extremes of every intermediate pixel value
    (EventConversion.cc:30-38), which depend on the event order; the device replays the window in order (ev_ordered_kernel)"""
    api = _api()
    ev = synth.make_events(6000, seed=77, w=240, h=180)
    cv = api.EvImConverter(0, 1, 100000, 346, 260)
    T = synth.rotation_tcw(np.array([0.5, -0.7, 1.5]) * float(ev["ts"][-1] - ev["ts"][0]))
    if mode == "nearest":
        u8 = cv.ev2im(ev, 240, 180, True, True); f = cv.ev2im(ev, 240, 180, True, False)
        rf, _, ru = O.ev_accumulate(ev, 240, 180, 1.0, mode=0, pol=True, normalize=True)
    elif mode == "gauss":
        f, u8 = cv.ev2im_gauss(ev, 240, 180, 1.0, True, True, both=True)
        rf, _, ru = O.ev_accumulate(ev, 240, 180, 1.0, mode=1, pol=True, normalize=True)
    elif mode == "se3":
        f, u8 = cv.ev2mci_gg_f(ev, K_ETHZ, T, 1.0, 240, 180, 1.0, True, True, both=True)
        rf, _, ru = O.ev_accumulate(ev, 240, 180, 1.0, mode=2, Tcw=T, depth=1.0, K=K_ETHZ, pol=True, normalize=True)
    else:
        f, u8 = cv.ev2mci_gg_f_2d(ev, K_ETHZ, [0.02, 1.0, -0.5], 240, 180, 1.0, True, True, both=True)
        rf, _, ru = O.ev_accumulate(ev, 240, 180, 1.0, mode=3, K=K_ETHZ, se2=[0.02, 1.0, -0.5], pol=True, normalize=True)
    if mode != "nearest":
        _close(f, rf)
    else:
        _close(f, rf)
    assert ru is not None and int(np.abs(u8.astype(int) - ru.astype(int)).max()) <= 1, "u8 frame with the reference's running extremes"
    # the final-frame extremes would NOT give this image: the running maximum is larger than the final one on this stream
    assert float(rf.max()) > 0 and float(rf.min()) < 0


def test_gauss_frame_of_a_very_large_window_stays_inside_the_tolerance():
    """one million events in one 240x180 window: the fixed-point scale of the shared-memory path is down to 2^13 here"""
    api = _api()
    ev = synth.make_events(1000000, seed=5, w=240, h=180)
    cv = api.EvImConverter(0, 1, 1 << 20, 240, 180)
    img = cv.ev2im_gauss(ev, 240, 180, 1.0, False, False)
    ref, _, _ = O.ev_accumulate(ev, 240, 180, 1.0, mode=1)
    _close(img, ref)
