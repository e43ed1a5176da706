"""Seeded cases and the HOST side of the keyframe-side searches of local mapping / loop closing -- ORBmatcher::SearchByProjection(KeyFrame*,
cv::Mat Scw, ...) (src/ORBmatcher.cc:480-593, :595-712), ORBmatcher::Fuse (:1407-1617, :1619-1741), ORBmatcher::SearchBySim3 (:1743-1967).

What runs on the device is the matching core (eorb_guided_search_windows: window lookup, level filter, reprojection gate, best
candidate, vpMatched blocking).  What the reference does around it with MapPoint / KeyFrame accessors -- the gates that decide whether a
point gets a window at all, and the map updates that follow a match -- stays on the host, in the order of the points; `host_windows` and
the `compose_*` functions below restate that host part (numpy float32, same operation order), so that
        reference function (libref, the reference's own bodies)  ==  host part + matching core (oracle or CUDA)
can be asserted.  Poses are the identity (see oracle/ref_guided_kf_api.cc), camera-frame point = world point."""
import os

import numpy as np

from eorb_slam_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AREA_QUERY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("r", "<f4"), ("min_level", "<i4"), ("max_level", "<i4")])
TH_LOW, TH_HIGH = 50, 100
F32 = np.float32


def _points(rng, x3, level_from, nlevels, bad_frac=0.05):
    """pt8 = pos[3] | normal[3] | minDist | maxDist per point; most pass the distance / viewing-angle gates, some fail each"""
    n = len(x3)
    pos = np.ascontiguousarray(x3, F32)
    dist = np.sqrt((pos.astype(np.float64) ** 2).sum(1)).astype(F32)
    nrm = (pos / np.maximum(dist, F32(1e-6))[:, None]).astype(F32)
    turn = rng.random(n) < 0.08                       # viewing angle above 60 degrees: PO . Pn < 0.5 * dist
    nrm[turn] *= F32(0.3)
    nrm[rng.random(n) < 0.03] *= F32(-1.0)
    mn = (dist * F32(0.5)).astype(F32); mx = (dist * F32(2.0)).astype(F32)
    out = rng.random(n) < 0.05
    mn[out] = dist[out] * F32(1.5)                    # closer than the scale-invariance region
    out = rng.random(n) < 0.05
    mx[out] = dist[out] * F32(0.7)
    pt8 = np.concatenate([pos, nrm, mn[:, None], mx[:, None]], 1).astype(F32)
    level = np.clip(level_from + rng.integers(-1, 2, n), 0, nlevels - 1).astype(np.int32)
    flags = (rng.random(n) < bad_frac).astype(np.uint8)
    return pt8, level, flags


def _bounds(seed):
    # every other case uses the bounds of a distorted camera: the Frame bins keypoints with the floats, the KeyFrame looks windows up with the
    # truncated ints (include/KeyFrame.h:529)
    return np.array([0, 0, 752, 480], F32) if seed % 2 == 0 else np.array([-12.7, -9.3, 764.2, 489.6], F32)


def kf_cases(nseeds=6):
    """(key, kind, overload, case dict)"""
    for seed in range(nseeds):
        rng = np.random.default_rng(7000 + seed)
        c = synth.make_projection_case(600, 640, 500 + seed)
        nl = len(c["scale_factors"])
        base = dict(kps2=c["kps2"], desc2=c["desc2"], bounds=_bounds(seed), K=c["K"], scale_factors=c["scale_factors"])
        pt8, level, flags = _points(rng, c["x3Dc"], c["kps1"]["octave"], nl)
        n1, n2 = len(level), len(c["kps2"])
        # SearchByProjection(KeyFrame*, Scw, ...): some slots hold a foreign point, some hold one of the candidate points (already found)
        held = np.full(n2, -1, np.int32)
        held[rng.random(n2) < 0.15] = -2
        own = rng.choice(n2, 30, replace=False)
        held[own] = rng.choice(n1, 30, replace=False)
        for ov in (0, 1):
            yield ("projkf%d_s%d" % (ov, seed), "projkf", ov,
                   dict(base, pt8=pt8, level1=level, flags1=flags, descMP=c["descMP"], held_id2=held, th=[3, 6, 10][seed % 3], ratio=[1.0, 0.8, 1.5][seed % 3]))
        # Fuse: occupied slots (some bad), observation counts on both sides, NULL entries, points already in the keyframe
        occ = np.zeros(n2, np.uint8)
        occ[rng.random(n2) < 0.4] = 1
        occ[rng.random(n2) < 0.05] = 2
        sig2 = (c["scale_factors"].astype(F32) ** 2).astype(F32)
        fz = dict(base, pt8=pt8, level1=level, descMP=c["descMP"], obs1=rng.integers(0, 9, n1).astype(np.int32),
                  flags1=(flags | ((rng.random(n1) < 0.05).astype(np.uint8) << 1)).astype(np.uint8), present1=(rng.random(n1) > 0.04).astype(np.uint8),
                  occupied2=occ, occ_obs2=rng.integers(0, 9, n2).astype(np.int32), inv_sigma2=(F32(1.0) / sig2).astype(F32),
                  mbf=F32(40.0), th=[3.0, 4.0, 2.5][seed % 3])
        ur2 = np.where(rng.random(n2) < 0.5, c["kps2"]["x"] - rng.random(n2).astype(F32) * 20, -1).astype(F32)
        yield "fuse0_s%d" % seed, "fuse", 0, dict(fz, u_right2=ur2 if seed % 2 else None)
        yield "fuse1_s%d" % seed, "fuse", 1, dict(fz, u_right2=None, flags1=flags, present1=np.ones(n1, np.uint8))
        # SearchBySim3: two keyframes seeing the same points (identity relative pose), each point's descriptor = its keyframe's
        c2 = synth.make_projection_case(520, 520, 600 + seed, shift=(2.0, -1.0))
        z = rng.uniform(1.0, 10.0, 520).astype(F32)
        K = c2["K"]

        def lift(k, sx, sy):
            return np.stack([(k["x"] + F32(sx) - K[2]) / K[0] * z[:len(k)], (k["y"] + F32(sy) - K[3]) / K[1] * z[:len(k)], z[:len(k)]], 1).astype(F32)
        p1, l1, f1 = _points(rng, lift(c2["kps1"], 2.0, -1.0), c2["kps1"]["octave"], nl)    # a point of KF1 projects onto its counterpart in KF2
        p2, l2, f2 = _points(rng, lift(c2["kps2"], -2.0, 1.0), c2["kps2"]["octave"], nl)
        m_in = np.full(520, -1, np.int32)
        m_in[rng.random(520) < 0.1] = -2
        pick = rng.choice(520, 25, replace=False)
        m_in[pick] = rng.choice(520, 25, replace=False)
        yield ("sim3_s%d" % seed, "sim3", 0,
               dict(kps1=c2["kps1"], desc1=c2["descMP"], kps2=c2["kps2"], desc2=c2["desc2"], bounds=_bounds(seed + 1), K=K, scale_factors=c2["scale_factors"],
                    pt8_1=p1, level_1=l1, flags_1=f1, present_1=(rng.random(520) > 0.1).astype(np.uint8),
                    pt8_2=p2, level_2=l2, flags_2=f2, present_2=(rng.random(520) > 0.1).astype(np.uint8), matched12_in=m_in, th=[7.5, 10.0][seed % 2]))


# ----------------------------------------------------------------------------- the host part (the reference's own code around the core)
def host_windows(pt8, level, skip, bounds, K, scale, th, proj_form, view_gate=True, mbf=None):
    """per point: the gates of ORBmatcher.cc:504-552 (= :621-667, :1448-1504, :1646-1696, :1790-1828) and the window.  proj_form 0 =
    pCamera->project(cv::Point3f) (Pinhole.cpp:30-33: fx * x / z + cx), 1 = invz = 1 / z; fx * (x * invz) + cx.  Returns (queries, ur)."""
    n = len(level)
    q = np.zeros(n, AREA_QUERY_DTYPE)
    q["r"] = -1.0; q["max_level"] = -1
    ur = np.zeros(n, F32)
    fx, fy, cx, cy = (F32(v) for v in K)
    bi = [int(v) for v in np.asarray(bounds, F32)]           # KeyFrame::mnMinX ... are ints (truncation toward zero)
    for i in range(n):
        if skip[i]:
            continue
        x, y, z = (F32(v) for v in pt8[i, :3])
        if z < 0.0:
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            invz = F32(1.0) / z
            if proj_form == 0:
                u = F32(F32(fx * x) / z) + cx; v = F32(F32(fy * y) / z) + cy
            else:
                u = F32(fx * F32(x * invz)) + cx; v = F32(fy * F32(y * invz)) + cy
        if not (u >= bi[0] and u < bi[2] and v >= bi[1] and v < bi[3]):          # KeyFrame::IsInImage (KeyFrame.cc:919-922)
            continue
        po = pt8[i, :3].astype(np.float64)
        dist = F32(np.sqrt((po * po).sum()))                                     # cv::norm: double accumulation, one rounding
        if dist < pt8[i, 6] or dist > pt8[i, 7]:
            continue
        if view_gate and float((po * pt8[i, 3:6].astype(np.float64)).sum()) < 0.5 * float(dist):
            continue
        lv = int(level[i])
        q[i] = (u, v, F32(F32(th) * F32(scale[lv])), lv - 1, lv)
        if mbf is not None:
            ur[i] = u - F32(F32(mbf) * invz)
    return q, ur


def _qmin(bounds):
    b = np.asarray(bounds, F32)
    return np.array([int(b[0]), int(b[1])], F32)


def compose_projkf(search, c, overload):
    n1 = len(c["level1"])
    already = np.zeros(n1, bool)
    hid = np.asarray(c["held_id2"])
    already[hid[hid >= 0]] = True                                                # spAlreadyFound (:496-497)
    skip = (np.asarray(c["flags1"]) & 1).astype(bool) | already
    q, _ = host_windows(c["pt8"], c["level1"], skip, c["bounds"], c["K"], c["scale_factors"], int(c["th"]), overload)
    th_high = int(np.floor(F32(TH_LOW) * F32(c["ratio"])))                       # bestDist <= TH_LOW * ratioHamming (int <= float)
    nm, _, _, m2 = search(q, None, c["descMP"], c["kps2"], c["desc2"], (hid != -1).astype(np.uint8), None, c["bounds"], query_min_xy=_qmin(c["bounds"]),
                          blocking=True, th_high=th_high)
    return nm, m2


def compose_fuse(search, c, overload):
    fl = np.asarray(c["flags1"])
    n1 = len(c["level1"])
    static_skip = np.asarray(c["present1"]) == 0
    # isBad() / IsInKeyFrame() are looked at when the point's turn comes (:1450-1459); the stand-ins never change them, so they are static here
    skip = static_skip | (fl & 1).astype(bool) | ((fl & 2).astype(bool) if overload == 0 else False)
    q, ur = host_windows(c["pt8"], c["level1"], skip, c["bounds"], c["K"], c["scale_factors"], c["th"], 0, mbf=c["mbf"] if overload == 0 else None)
    kw = dict(inv_level_sigma2=c["inv_sigma2"]) if overload == 0 else {}
    _, bi, _, _ = search(q, ur if overload == 0 else None, c["descMP"], c["kps2"], c["desc2"], None, c.get("u_right2") if overload == 0 else None, c["bounds"],
                         query_min_xy=_qmin(c["bounds"]), blocking=False, th_high=TH_LOW, **kw)
    # the map updates, in the order of the points (:1583-1602, :1727-1740)
    occ = np.asarray(c["occupied2"]); occ_obs = np.asarray(c["occ_obs2"]); obs1 = np.asarray(c["obs1"])
    slot = {int(s): -(int(s) + 2) for s in np.nonzero(occ)[0]}                    # id of the point sitting in a slot
    ev, nfused = [], 0
    for i in range(n1):
        b = int(bi[i])
        if b < 0:
            continue
        if b in slot:
            o = slot[b]
            o_bad = (occ[-(o + 2)] == 2) if o < 0 else bool(fl[o] & 1)
            o_obs = int(occ_obs[-(o + 2)]) if o < 0 else int(obs1[o])
            if not o_bad:
                if overload == 0:
                    ev.append((1, i, o) if o_obs > int(obs1[i]) else (1, o, i))
                else:
                    ev.append((2, i, o))
        else:
            slot[b] = i
            ev.append((0, i, b))
        nfused += 1
    return nfused, np.array(ev, np.int32).reshape(-1, 3)


def compose_sim3(search, c):
    n1, n2 = len(c["kps1"]), len(c["kps2"])
    m_in = np.asarray(c["matched12_in"])
    am1 = m_in != -1
    am2 = np.zeros(n2, bool)
    am2[m_in[m_in >= 0]] = True                                                  # GetIndexInKeyFrame(pKF2) of the matched points (:1771-1782)
    sk1 = (np.asarray(c["present_1"]) == 0) | am1 | (np.asarray(c["flags_1"]) & 1).astype(bool)
    sk2 = (np.asarray(c["present_2"]) == 0) | am2 | (np.asarray(c["flags_2"]) & 1).astype(bool)
    q1, _ = host_windows(c["pt8_1"], c["level_1"], sk1, c["bounds"], c["K"], c["scale_factors"], c["th"], 1, view_gate=False)
    q2, _ = host_windows(c["pt8_2"], c["level_2"], sk2, c["bounds"], c["K"], c["scale_factors"], c["th"], 1, view_gate=False)
    qm = _qmin(c["bounds"])
    _, b1, _, _ = search(q1, None, c["desc1"], c["kps2"], c["desc2"], None, None, c["bounds"], query_min_xy=qm, blocking=False, th_high=TH_HIGH)
    _, b2, _, _ = search(q2, None, c["desc2"], c["kps1"], c["desc1"], None, None, c["bounds"], query_min_xy=qm, blocking=False, th_high=TH_HIGH)
    m12 = m_in.copy().astype(np.int32)
    nfound = 0
    for i1 in range(n1):                                                         # agreement (:1944-1958)
        i2 = int(b1[i1])
        if i2 >= 0 and int(b2[i2]) == i1:
            m12[i1] = i2
            nfound += 1
    return nfound, m12


def run_ref(R, kind, overload, c):
    if kind == "projkf":
        return R.search_by_projection_kf(c, overload)
    if kind == "fuse":
        return R.fuse(c, overload)
    return R.search_by_sim3(c)


def run_composed(search, kind, overload, c):
    if kind == "projkf":
        return compose_projkf(search, c, overload)
    if kind == "fuse":
        return compose_fuse(search, c, overload)
    return compose_sim3(search, c)


def kf_golden():
    return np.load(os.path.join(GOLDEN, "ref_guided_kf.npz"))


# ----------------------------------------------------------------------------- SearchForTriangulation (ORBmatcher.cc:975-1214)
def tri_cases(nseeds=4):
    """(key, args of ref_lib.search_for_triangulation, kwargs).  Geometries: a sideways translation along the image flow (epipole at infinity,
    true matches lie near their epipolar lines), and a forward motion that puts the epipole inside the image (epipole-distance gate, most
    matches off their lines).  Variants: right-image columns, bOnlyStereo, bCoarse, rotation check off."""
    import ref_cases as RC
    K = np.array([458.654, 457.296, 367.215, 248.375], F32)
    for seed in range(nseeds):
        rng = np.random.default_rng(8000 + seed)
        k1, d1, _, fv1, k2, d2, fv2 = RC.bow_case(600, 640, 700 + seed, levelsup=1 + seed % 3, max_flips=30)
        nl = 8
        sc = np.ones(nl, F32)
        for i in range(1, nl):
            sc[i] = F32(np.float64(sc[i - 1]) * np.float64(F32(1.2)))
        sg = (sc * sc).astype(F32)
        h1 = (rng.random(len(k1)) < 0.3).astype(np.uint8); h2 = (rng.random(len(k2)) < 0.3).astype(np.uint8)
        u1 = np.where(rng.random(len(k1)) < 0.4, k1["x"] - 5, -1).astype(F32); u2 = np.where(rng.random(len(k2)) < 0.4, k2["x"] - 5, -1).astype(F32)
        geos = {"side": (np.zeros(3, F32), np.array([-0.07, 0.04, 0.0], F32)), "fwd": (np.array([0.01, 0.0, 0.0], F32), np.array([0.02, -0.03, -1.0], F32))}
        for gname, (t1, t2) in geos.items():
            for var, (ur, only, coarse, ori) in {"mono": (False, False, False, True), "st": (True, False, False, True), "only": (True, True, False, bool(seed % 2)),
                                                 "coarse": (False, False, True, True), "noori": (False, False, False, False)}.items():
                a = (k1, d1, h1, u1 if ur else None, fv1, k2, d2, h2, u2 if ur else None, fv2, K, t1, t2, sc, sg)
                yield "tri_%s_%s_s%d" % (gname, var, seed), a, dict(only_stereo=only, coarse=coarse, check_ori=ori)


def tri_compose(search, a, kw, F12, ep):
    """the device-path call for one case: flags from the map-point slots / right columns, F12 and the epipole as the caller forms them"""
    import oracle_lib as O
    k1, d1, h1, u1, fv1, k2, d2, h2, u2, fv2, K, t1, t2, sc, sg = a
    f1 = O.triangulation_flags(h1, u1, kw["only_stereo"]); f2 = O.triangulation_flags(h2, u2, kw["only_stereo"])
    return search(k1, d1, f1, fv1, k2, d2, f2, fv2, F12, ep, sc, sg, kw["coarse"], kw["check_ori"])


def tri_golden():
    return np.load(os.path.join(GOLDEN, "ref_triangulation.npz"))

