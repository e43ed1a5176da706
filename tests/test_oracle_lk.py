"""Pins the LK oracle (oracle/lk_oracle.cc) against OpenCV itself (cv2 wheel): pyrDown bit-exactly, the Scharr
derivative against cv2.Scharr, and the tracked points against cv2.calcOpticalFlowPyrLK within the freedom OpenCV's own
SIMD summation order leaves."""
import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth

cv2 = pytest.importorskip("cv2")


def _pair(seed, w, h, shift=(1.7, -0.9)):
    """a textured frame and a sub-pixel translated + slightly rotated copy"""
    img = synth.make_frame(seed, w, h)
    M = cv2.getRotationMatrix2D((w / 2, h / 2), 0.4, 1.0)
    M[0, 2] += shift[0]; M[1, 2] += shift[1]
    nxt = cv2.warpAffine(img, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    return img, nxt


@pytest.mark.parametrize("wh", [(240, 180), (346, 260), (97, 131), (33, 47), (752, 480)])
def test_pyrdown_equals_cv2(wh):
    w, h = wh
    img = np.random.default_rng(w + h).integers(0, 256, (h, w), dtype=np.uint8)
    assert np.array_equal(O.pyrdown(img), cv2.pyrDown(img))


@pytest.mark.parametrize("wh", [(240, 180), (61, 37), (2, 9)])
def test_scharr_deriv_equals_cv2(wh):
    w, h = wh
    img = np.random.default_rng(w * 7 + h).integers(0, 256, (h, w), dtype=np.uint8)
    d = O.scharr_deriv(img)
    gx = cv2.Scharr(img, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101)
    gy = cv2.Scharr(img, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(d[..., 0], gx) and np.array_equal(d[..., 1], gy)


@pytest.mark.parametrize("cfg", [dict(w=240, h=180, win=23, lv=1), dict(w=346, h=260, win=23, lv=1), dict(w=240, h=180, win=15, lv=3),
                                 dict(w=752, h=480, win=21, lv=2)])
def test_lk_matches_cv2_within_summation_order_freedom(cfg):
    w, h, win, lv = cfg["w"], cfg["h"], cfg["win"], cfg["lv"]
    img, nxt = _pair(5 + w, w, h)
    rng = np.random.default_rng(w * 3 + win)
    corners = cv2.goodFeaturesToTrack(img, 300, 0.01, 5).reshape(-1, 2)
    # plus points near / beyond the border and in flat regions
    extra = np.stack([rng.uniform(-5, w + 5, 60), rng.uniform(-5, h + 5, 60)], 1).astype(np.float32)
    pts = np.concatenate([corners, extra]).astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 10, 0.03)
    ref_p, ref_s, ref_e = cv2.calcOpticalFlowPyrLK(img, nxt, pts.reshape(-1, 1, 2), None, winSize=(win, win), maxLevel=lv, criteria=crit)
    ref_p = ref_p.reshape(-1, 2); ref_s = ref_s.reshape(-1); ref_e = ref_e.reshape(-1)
    got_p, got_s, got_e, _ = O.lk_track(img, nxt, pts, None, win, lv, 10, 0.03)
    same = got_s == ref_s
    assert same.mean() >= 0.99, "status differs on %d points" % int((~same).sum())
    ok = same & (ref_s == 1)
    d = np.abs(got_p[ok] - ref_p[ok]).max(axis=1)
    assert (d <= 0.02).mean() >= 0.99 and np.median(d) <= 1e-3, (float(np.median(d)), float(d.max()))
    de = np.abs(got_e[ok] - ref_e[ok])
    assert (de <= 0.05).mean() >= 0.99

    # OPTFLOW_USE_INITIAL_FLOW (the form ELK_Tracker uses once it has tracked points, KLT_Tracker.cpp:63-65, 86-88)
    init = (pts + rng.uniform(-1.5, 1.5, pts.shape)).astype(np.float32)
    ref_p2, ref_s2, _ = cv2.calcOpticalFlowPyrLK(img, nxt, pts.reshape(-1, 1, 2), init.reshape(-1, 1, 2).copy(), winSize=(win, win),
                                                  maxLevel=lv, criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    got_p2, got_s2, _, _ = O.lk_track(img, nxt, pts, init, win, lv, 10, 0.03)
    ref_p2 = ref_p2.reshape(-1, 2); ref_s2 = ref_s2.reshape(-1)
    same2 = got_s2 == ref_s2
    assert same2.mean() >= 0.99
    ok2 = same2 & (ref_s2 == 1)
    d2 = np.abs(got_p2[ok2] - ref_p2[ok2]).max(axis=1)
    assert (d2 <= 0.02).mean() >= 0.99


# ---- contrast metric of event frames: oracle vs a cv2.meanStdDev restatement of the reference loop
def _focus_cv2(img, what, avg, patch=30):
    h, w = img.shape
    if what == O.FOCUS_GLOBAL_STD:
        return np.float32(cv2.meanStdDev(img)[1][0, 0])
    vals = []
    acc = np.float32(0)
    for i in range(0, h, patch):
        for j in range(0, w, patch):
            m, s = cv2.meanStdDev(img[i:min(i + patch, h), j:min(j + patch, w)])
            v = np.float32(m[0, 0] if what == O.FOCUS_LOCAL_MEAN else s[0, 0])
            acc = np.float32(acc + v); vals.append(v)
    if avg:
        return np.float32(acc / np.float32(len(vals)))
    return sorted(vals)[len(vals) // 2]


@pytest.mark.parametrize("wh", [(240, 180), (346, 260), (61, 47), (30, 30), (29, 95)])
def test_image_focus_equals_cv2_mean_std_dev(wh):
    w, h = wh
    if w >= 200:
        ev = synth.make_events(20000, seed=w, w=w, h=h)
        img, _, _ = O.ev_accumulate(ev, w, h, 1.0, mode=1)
    else:
        img = (np.random.default_rng(w * h).random((h, w)) ** 3 * 4).astype(np.float32)
    for what in (O.FOCUS_LOCAL_STD, O.FOCUS_GLOBAL_STD, O.FOCUS_LOCAL_MEAN):
        for avg in (True, False):
            got, exp = O.image_focus(img, what, avg), float(_focus_cv2(img, what, avg))
            assert abs(got - exp) <= 2e-6 * max(abs(exp), 1e-3), (what, avg, got, exp)


def test_mci_jacobian_oracle_agrees_with_finite_differences_of_the_contrast():
    """ev2mci_gg_f_jac (EventConversion.cc:533-662) is the gradient of mean(I^2) of the motion-compensated frame w.r.t.
    the window motion; for the translation part the reference's per-event model (t_k = t * rate) is exactly linear, so a
    central difference of the oracle's own E4 frames must reproduce jac[3:6] up to the frame's float discretisation."""
    w, h = 240, 180
    ev = synth.make_events(6000, seed=3, w=w, h=h)
    K = np.array((199.09, 198.83, 132.19, 110.71), np.float32)
    dt = float(ev["ts"][-1] - ev["ts"][0])
    T = synth.rotation_tcw(np.array([0.5, -0.7, 1.5]) * dt).astype(np.float64)
    t0 = np.array([0.02, -0.01, 0.03])

    def contrast(t):
        Tm = T.copy(); Tm[:3, 3] = t
        img, _, _ = O.ev_accumulate(ev, w, h, 1.0, mode=2, Tcw=Tm.astype(np.float32), depth=1.0, K=K)
        return float((img.astype(np.float64) ** 2).mean())

    jac = O.ev_mci_jac(ev, w, h, 1.0, T[:3, :3], t0, 1.0, K, False, True)
    step = 1e-3
    fd = np.array([(contrast(t0 + step * np.eye(3)[k]) - contrast(t0 - step * np.eye(3)[k])) / (2 * step) for k in range(3)])
    big = np.abs(fd) > 0.1
    assert big.any() and np.all(np.sign(jac[3:][big]) == np.sign(fd[big]))
    assert np.all(np.abs(jac[3:][big] - fd[big]) <= 0.5 * np.abs(fd[big]))
