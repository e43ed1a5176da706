#!/usr/bin/env python3
"""Pin the CPU oracle against cv2 4.13.0 and write the golden fixtures in this directory.

The reference (m-dayani/EORB_SLAM) has no tests or golden vectors and cannot be compiled in this
environment (needs OpenCV 3.4.1 C++ headers); its arithmetic lives in OpenCV primitives.  This script
(run in the authoring container, where the Python cv2 wheel exists) therefore
  1. checks every OpenCV primitive restated in oracle/ bit-for-bit against cv2 on random inputs,
  2. runs an INDEPENDENT cv2-assisted Python restatement of ORBextractor::operator()
     (reference src/ORBextractor.cc:1092-1176; cv2 does resize/border/FAST/blur/fastAtan2, Python lists
     emulate DistributeOctTree's std::list) and demands identical keypoints+descriptors from the oracle,
  3. stores small input/output vectors as .npz so the GPU box (no cv2 guarantee, no /root/reference)
     can re-check both the oracle and the CUDA path against them.

Usage:  python tests/golden/make_golden.py        (asserts, then rewrites tests/golden/*.npz)
"""
import math
import os
import sys
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
from eorb_slam_b200 import synth  # noqa: E402

cv2.setNumThreads(1)
PATTERN = np.array([int(v) for v in "".join(
    l for l in open(os.path.join(ROOT, "oracle", "brief_pattern_31.inc")) if not l.startswith("//")
).replace("\n", "").split(",") if v.strip()], np.int32).reshape(512, 2)


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


# --------------------------------------------------------------------------- primitive pins
def pin_primitives():
    rng = np.random.default_rng(1234)
    out = {}
    # resize: every pyramid transition for the four (size, factor, levels) sets in the YAMLs + odd shapes
    n_resize = 0
    for (w, h, s, nl) in [(752, 480, 1.2, 8), (240, 180, 1.2, 4), (346, 260, 1.26, 6), (346, 260, 1.1, 16), (97, 131, 1.2, 5)]:
        orc = O.OrbOracle(1000, s, nl, 20, 7, 19, w, h)
        _, inv, _, _ = orc.scale_factors()
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        prev = img
        for l in range(1, nl):
            dw = O.lib().orc_cv_round_f(np.float32(w) * inv[l]); dh = O.lib().orc_cv_round_f(np.float32(h) * inv[l])
            ref = cv2.resize(prev, (dw, dh), interpolation=cv2.INTER_LINEAR)
            got = O.resize_linear(prev, dw, dh)
            assert np.array_equal(ref, got), ("resize", w, h, s, l, int((ref != got).sum()))
            prev = ref
            n_resize += 1
    small = rng.integers(0, 256, (131, 97), dtype=np.uint8)
    out["small"] = small
    out["small_resize_81x109"] = cv2.resize(small, (81, 109), interpolation=cv2.INTER_LINEAR)
    out["small_resize_50x77"] = cv2.resize(small, (50, 77), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(out["small_resize_81x109"], O.resize_linear(small, 81, 109))
    assert np.array_equal(out["small_resize_50x77"], O.resize_linear(small, 50, 77))
    # border
    for b in (9, 15, 19):
        ref = cv2.copyMakeBorder(small, b, b, b, b, cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, O.border_reflect101(small, b))
    out["small_border19"] = cv2.copyMakeBorder(small, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    # blur
    for shape in [(131, 97), (480, 752), (7, 9), (5, 5), (33, 64)]:
        im = rng.integers(0, 256, shape, dtype=np.uint8)
        ref = cv2.GaussianBlur(im, (5, 5), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, O.gauss5(im)), ("blur", shape)
    smooth = cv2.GaussianBlur(synth.make_frame(3), (5, 5), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(smooth, O.gauss5(synth.make_frame(3)))
    out["small_blur"] = cv2.GaussianBlur(small, (5, 5), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    # FAST: random noise tiles, textured tiles, thresholds 0/1/7/10/20, cell-sized ROIs
    n_fast = 0
    frame = synth.make_frame(5)
    for t in (0, 1, 7, 10, 20, 40):
        det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True)
        for trial in range(6):
            if trial % 2 == 0:
                tile = rng.integers(0, 256, (int(rng.integers(7, 48)), int(rng.integers(7, 48))), dtype=np.uint8)
            else:
                y0 = int(rng.integers(0, 430)); x0 = int(rng.integers(0, 700))
                tile = np.ascontiguousarray(frame[y0:y0 + int(rng.integers(8, 47)), x0:x0 + int(rng.integers(8, 47))])
            kps = det.detect(tile)
            ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps]
            xs, ys, sc = O.fast(tile, t, True)
            got = list(zip(xs.tolist(), ys.tolist(), sc.tolist()))
            assert ref == got, ("fast", t, trial, len(ref), len(got))
            n_fast += 1
    det = cv2.FastFeatureDetector_create(threshold=7, nonmaxSuppression=False)
    tile = np.ascontiguousarray(frame[100:140, 200:242])
    ref = [(int(k.pt[0]), int(k.pt[1])) for k in det.detect(tile)]
    xs, ys, sc = O.fast(tile, 7, False)
    assert ref == list(zip(xs.tolist(), ys.tolist()))
    tile = np.ascontiguousarray(frame[60:102, 300:346])
    out["fast_tile"] = tile
    for t in (0, 7, 20):
        det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True)
        k = det.detect(tile)
        out["fast_tile_t%d" % t] = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k], np.int32).reshape(-1, 3)
    # fastAtan2
    ys_ = rng.integers(-3000000, 3000000, 4000).astype(np.float32)
    xs_ = rng.integers(-3000000, 3000000, 4000).astype(np.float32)
    ys_[:8] = [0, 0, 1, -1, 5, -5, 0, 7]; xs_[:8] = [0, 1, 0, 0, 5, 5, -3, -7]
    ref = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys_, xs_)], np.float32)
    got = np.array([O.fast_atan2(float(y), float(x)) for y, x in zip(ys_, xs_)], np.float32)
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32)), "fastAtan2"
    out["atan_y"] = ys_[:512]; out["atan_x"] = xs_[:512]; out["atan_out"] = ref[:512]
    # cvRound ties
    for v in (0.5, 1.5, 2.5, -0.5, -1.5, 3.4999, 1e6 + 0.5):
        assert O.lib().orc_cv_round_f(np.float32(v)) == int(np.rint(np.float32(v)))
    # brute-force best-2 incl. ties vs cv2.BFMatcher
    db = rng.integers(0, 256, (600, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    q[:20] = db[rng.integers(0, 600, 20)]
    db[500:520] = db[100:120]                      # exact duplicates -> distance ties
    q[20:30] = db[105:115]
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    knn = bf.knnMatch(q, db, 2)
    m = O.hamming_best2(q, db, th=50, ratio=0.7)
    for i, (a, b) in enumerate(knn):
        assert (int(a.distance), a.trainIdx, int(b.distance)) == (m[i]["best_dist"], m[i]["best_idx"], m[i]["second_dist"]), i
        assert m[i]["accepted"] == int(a.distance <= 50 and np.float32(a.distance) < np.float32(0.7) * np.float32(b.distance))
        assert O.descriptor_distance(q[i], db[a.trainIdx]) == int(a.distance)
    out["bf_db"] = db; out["bf_q"] = q
    out["bf_out"] = np.array([(int(a.distance), a.trainIdx, int(b.distance)) for a, b in knn], np.int32)
    # normalisation
    f = (rng.random((180, 240)).astype(np.float32) ** 3) * np.float32(7.3)
    ref = cv2.normalize(f, None, 255, 0, cv2.NORM_MINMAX, cv2.CV_8UC1)
    got = O.normalize_minmax_u8(f)
    assert np.abs(ref.astype(int) - got.astype(int)).max() <= 1, "normalize"
    out["norm_mismatch_px"] = np.array([int((ref != got).sum())])
    mx = float(f.max()); a = np.float32(255.0) / np.float32(mx)
    ref2 = cv2.convertScaleAbs(f, alpha=float(a), beta=0.0)
    got2 = np.empty_like(ref2)
    O.lib().orc_normalize_convert_u8(f.ctypes.data, f.size, mx, 0.0, got2.ctypes.data)
    assert np.abs(ref2.astype(int) - got2.astype(int)).max() <= 1, "convertTo"
    print("primitives pinned: %d resize transitions, %d FAST tiles, 4000 atan2, knn ties, blur, border, normalize(%d px off by 1)"
          % (n_resize, n_fast, int((ref != got).sum())))
    np.savez_compressed(os.path.join(HERE, "prims.npz"), **out)


# --------------------------------------------------------------------------- independent pipeline
class PyNode:
    __slots__ = ("keys", "UL", "UR", "BL", "BR", "nomore", "seq", "alive")

    def __init__(self):
        self.keys = []; self.nomore = False; self.seq = 0; self.alive = True


def py_divide(n):
    halfX = int(math.ceil(np.float32(n.UR[0] - n.UL[0]) / 2)); halfY = int(math.ceil(np.float32(n.BR[1] - n.UL[1]) / 2))
    c = [PyNode() for _ in range(4)]
    c[0].UL = n.UL; c[0].UR = (n.UL[0] + halfX, n.UL[1]); c[0].BL = (n.UL[0], n.UL[1] + halfY); c[0].BR = (n.UL[0] + halfX, n.UL[1] + halfY)
    c[1].UL = c[0].UR; c[1].UR = n.UR; c[1].BL = c[0].BR; c[1].BR = (n.UR[0], n.UL[1] + halfY)
    c[2].UL = c[0].BL; c[2].UR = c[0].BR; c[2].BL = n.BL; c[2].BR = (c[0].BR[0], n.BL[1])
    c[3].UL = c[2].UR; c[3].UR = c[1].BR; c[3].BL = c[2].BR; c[3].BR = n.BR
    for k in n.keys:
        if k[0] < c[0].UR[0]:
            (c[0] if k[1] < c[0].BR[1] else c[2]).keys.append(k)
        elif k[1] < c[0].BR[1]:
            c[1].keys.append(k)
        else:
            c[3].keys.append(k)
    for x in c:
        x.nomore = len(x.keys) == 1
    return c


def py_octree(keys, minX, maxX, minY, maxY, N):
    """keys: list of (x, y, resp, idx).  Python-list emulation of DistributeOctTree (:558-782).
    `nodes` is the std::list (index 0 = front)."""
    nIni = int(np.round(np.float32(maxX - minX) / np.float32(maxY - minY)))
    hX = np.float32(maxX - minX) / np.float32(nIni)
    nodes = []
    seq = 0
    for i in range(nIni):
        n = PyNode()
        n.UL = (int(hX * np.float32(i)), 0); n.UR = (int(hX * np.float32(i + 1)), 0)
        n.BL = (n.UL[0], maxY - minY); n.BR = (n.UR[0], maxY - minY)
        n.seq = seq; seq += 1
        nodes.append(n)
    for k in keys:
        nodes[int(np.float32(k[0]) / hX)].keys.append(k)
    nodes = [n for n in nodes if n.keys]
    for n in nodes:
        n.nomore = len(n.keys) == 1
    finish = False
    while not finish:
        prev_size = len(nodes)
        n_to_expand = 0
        vsize = []
        front = []
        rest = []
        for n in nodes:               # iteration order == list order; children are push_front-ed
            if n.nomore:
                rest.append(n); continue
            for c in py_divide(n):
                if c.keys:
                    c.seq = seq; seq += 1
                    front.insert(0, c)
                    if len(c.keys) > 1:
                        n_to_expand += 1; vsize.append(c)
        nodes = front + rest
        if len(nodes) >= N or len(nodes) == prev_size:
            finish = True
        elif len(nodes) + n_to_expand * 3 > N:
            while not finish:
                prev_size = len(nodes)
                prev = sorted(vsize, key=lambda c: (len(c.keys), c.seq))
                vsize = []
                for j in range(len(prev) - 1, -1, -1):
                    for c in py_divide(prev[j]):
                        if c.keys:
                            c.seq = seq; seq += 1
                            nodes.insert(0, c)
                            if len(c.keys) > 1:
                                vsize.append(c)
                    nodes.remove(prev[j])
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev_size:
                    finish = True
    res = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if k[2] > best[2]:
                best = k
        res.append(best)
    return res


def py_orb(img, nfeatures, scale_factor, nlevels, ini_th, min_th, edge, lapping=(0, 1000)):
    """cv2-assisted restatement of operator() (:1092-1176).  Returns (monoIndex, kps, desc, stage info)."""
    f32 = np.float32
    sf = float(f32(scale_factor))
    scale = [f32(1.0)]
    for i in range(1, nlevels):
        scale.append(f32(float(scale[-1]) * sf))
    inv = [f32(1.0) / s for s in scale]
    factor = f32(1.0 / sf)
    nd = f32(f32(nfeatures) * (f32(1) - factor)) / (f32(1) - f32(math.pow(float(factor), float(nlevels)))) if nlevels > 1 else f32(0)
    quota = []
    for l in range(nlevels - 1):
        quota.append(int(np.rint(nd))); nd = f32(nd * factor)
    quota.append(max(nfeatures - sum(quota), 0))
    umax = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    E = edge
    h, w = img.shape
    pyr = []
    for l in range(nlevels):
        sw = int(np.rint(f32(w) * inv[l])); sh = int(np.rint(f32(h) * inv[l]))
        lvl = img if l == 0 else cv2.resize(pyr[l - 1][E:-E, E:-E], (sw, sh), interpolation=cv2.INTER_LINEAR)
        pyr.append(cv2.copyMakeBorder(lvl, E, E, E, E, cv2.BORDER_REFLECT_101))
    det_ini = cv2.FastFeatureDetector_create(threshold=ini_th, nonmaxSuppression=True)
    det_min = cv2.FastFeatureDetector_create(threshold=min_th, nonmaxSuppression=True)
    all_kps = []
    info = {"ncand": [], "nfallback": 0, "quota": quota}
    for l in range(nlevels):
        lv = pyr[l][E:-E, E:-E]
        H, W = lv.shape
        minBX = E - 3; minBY = E - 3; maxBX = W - E + 3; maxBY = H - E + 3
        width = f32(maxBX - minBX); height = f32(maxBY - minBY)
        nCols = int(width / f32(30)); nRows = int(height / f32(30))
        cand = []
        if nCols > 0 and nRows > 0:
            wCell = int(math.ceil(width / f32(nCols))); hCell = int(math.ceil(height / f32(nRows)))
            for i in range(nRows):
                iniY = minBY + i * hCell; maxY = iniY + hCell + 6
                if iniY >= maxBY - 3:
                    continue
                maxY = min(maxY, maxBY)
                for j in range(nCols):
                    iniX = minBX + j * wCell; maxX = iniX + wCell + 6
                    if iniX >= maxBX - 3:
                        continue
                    maxX = min(maxX, maxBX)
                    roi = np.ascontiguousarray(lv[iniY:maxY, iniX:maxX])
                    k = det_ini.detect(roi)
                    if not k:
                        info["nfallback"] += 1
                        k = det_min.detect(roi)
                    for p in k:
                        cand.append((f32(p.pt[0] + j * wCell), f32(p.pt[1] + i * hCell), f32(p.response), len(cand)))
        info["ncand"].append(len(cand))
        sel = py_octree(cand, minBX, maxBX, minBY, maxBY, quota[l]) if (maxBX > minBX and maxBY > minBY) else []
        lk = []
        for (x, y, r, _) in sel:
            lk.append([f32(x + minBX), f32(y + minBY), f32(int(31 * float(scale[l]))), f32(-1), r, l])
        # orientation on the bordered level
        for k in lk:
            cx = int(np.rint(k[0])) + E; cy = int(np.rint(k[1])) + E
            m01 = 0; m10 = 0
            for v in range(-15, 16):
                d = umax[abs(v)]
                row = pyr[l][cy + v, cx - d:cx + d + 1].astype(np.int64)
                m10 += int((np.arange(-d, d + 1) * row).sum()); m01 += v * int(row.sum())
            k[3] = f32(cv2.fastAtan2(float(m01), float(m10)))
        all_kps.append(lk)
    nk = sum(len(k) for k in all_kps)
    kps = np.zeros(nk, O.KEYPOINT_DTYPE)
    desc = np.zeros((nk, 32), np.uint8)
    mono = 0; stereo = nk - 1
    factor_pi = f32(math.pi / 180.0)
    for l in range(nlevels):
        if not all_kps[l]:
            continue
        lv = np.ascontiguousarray(pyr[l][E:-E, E:-E])
        bl = cv2.GaussianBlur(lv, (5, 5), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        H, W = bl.shape
        for k in all_kps[l]:
            ang = f32(k[3] * factor_pi)
            a = f32(math.cos(float(ang))); b = f32(math.sin(float(ang)))
            cy = int(np.rint(k[1])); cx = int(np.rint(k[0]))
            px = PATTERN[:, 0].astype(f32); py = PATTERN[:, 1].astype(f32)
            rr = np.rint(px * b + py * a).astype(np.int64) + cy
            cc = np.rint(px * a - py * b).astype(np.int64) + cx
            rr = np.where(rr < 0, -rr, rr); rr = np.where(rr >= H, 2 * H - 2 - rr, rr)
            cc = np.where(cc < 0, -cc, cc); cc = np.where(cc >= W, 2 * W - 2 - cc, cc)
            vals = bl[rr, cc].astype(np.int32)
            bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
            d = np.packbits(bits, bitorder="little")
            x, y = k[0], k[1]
            if l != 0:
                x = f32(x * scale[l]); y = f32(y * scale[l])
            if x >= lapping[0] and x <= lapping[1]:
                idx = stereo; stereo -= 1
            else:
                idx = mono; mono += 1
            kps[idx] = (x, y, k[2], k[3], k[4], l, -1)
            desc[idx] = d
    info["pyr_crc"] = [crc(pyr[l][E:-E, E:-E]) for l in range(nlevels)]
    return mono, kps, desc, info


def pin_pipeline():
    cases = [
        ("cfg1_seed0", dict(seed=0, w=752, h=480, kind="textured"), dict(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, edge=19), (0, 1000)),
        ("cfg1_seed1_stereo", dict(seed=1, w=752, h=480, kind="textured"), dict(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, edge=19), (0, 0)),
        ("cfg1_flat", dict(seed=2, w=752, h=480, kind="flat"), dict(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, edge=19), (0, 1000)),
        ("mvsec_346x260", dict(seed=3, w=346, h=260, kind="textured"), dict(nfeatures=1000, scale_factor=1.26, nlevels=6, ini_th=10, min_th=1, edge=19), (0, 1000)),
        ("ethz_240x180_e9", dict(seed=4, w=240, h=180, kind="textured"), dict(nfeatures=1000, scale_factor=1.2, nlevels=4, ini_th=10, min_th=0, edge=9), (0, 1000)),
        ("ev_single_level", dict(seed=5, w=240, h=180, kind="textured"), dict(nfeatures=400, scale_factor=1.0, nlevels=1, ini_th=0, min_th=0, edge=9), (0, 1000)),
    ]
    for name, fkw, okw, lap in cases:
        img = synth.make_frame(**fkw)
        mono, kps, desc, info = py_orb(img, lapping=lap, **okw)
        orc = O.OrbOracle(okw["nfeatures"], okw["scale_factor"], okw["nlevels"], okw["ini_th"], okw["min_th"], okw["edge"], fkw["w"], fkw["h"])
        ret, okps, odesc = orc.extract(img, lap, True)
        assert list(orc.features_per_level()) == info["quota"], (name, orc.features_per_level(), info["quota"])
        assert [crc(orc.level(l)) for l in range(okw["nlevels"])] == info["pyr_crc"], name
        assert [len(orc.candidates(l)[0]) for l in range(okw["nlevels"])] == info["ncand"], (name, info["ncand"])
        assert orc.fallback_cells() == info["nfallback"], name
        assert ret == mono and len(okps) == len(kps), (name, ret, mono, len(okps), len(kps))
        assert okps.tobytes() == kps.tobytes(), name
        nbad = int((odesc != desc).any(axis=1).sum())
        assert nbad == 0, (name, "descriptor rows differ", nbad)
        print("pipeline pinned: %-20s kps=%4d mono=%4d cand=%s fallback_cells=%d" % (name, len(kps), mono, info["ncand"], info["nfallback"]))
        np.savez_compressed(os.path.join(HERE, "orb_%s.npz" % name), frame_kw=np.array(repr(fkw)), orb_kw=np.array(repr(okw)),
                            lapping=np.array(lap, np.int32), ret=np.array([mono], np.int32), kps=kps, desc=desc,
                            ncand=np.array(info["ncand"], np.int32), nfallback=np.array([info["nfallback"]], np.int32),
                            quota=np.array(info["quota"], np.int32), pyr_crc=np.array(info["pyr_crc"], np.uint32),
                            frame_crc=np.array([crc(img)], np.uint32))


# --------------------------------------------------------------------------- events
def py_gauss_splat(evxy, w, h, sigma):
    """float32 sequential restatement of ev2im_gauss (pol=false) in plain Python (small inputs only)."""
    f32 = np.float32
    img = np.zeros((h, w), f32)
    sig2 = f32(sigma) * f32(sigma)
    half = int(math.ceil(sigma * 3.0))
    norm = f32(2.0) * f32(math.pi) * sig2
    for X, Y in evxy:
        X = f32(X); Y = f32(Y)
        xi = int(math.floor(X)); yi = int(math.floor(Y))
        xr = f32(X - f32(xi)); yr = f32(Y - f32(yi))
        for i in range(-half, half + 1):
            for j in range(-half, half + 1):
                xn = xi + i; yn = yi + j
                if not (0 <= xn < w and 0 <= yn < h):
                    continue
                dx = f32(f32(i) - xr); dy = f32(f32(j) - yr)
                dd = f32(f32(dx * dx) + f32(dy * dy)) / f32(f32(2.0) * sig2)
                val = f32(f32(math.exp(-float(dd))) / norm)
                img[yn, xn] = f32(img[yn, xn] + val)
    return img


def pin_events():
    ev = synth.make_events(2000, seed=11, w=240, h=180)
    img, (mn, mx), u8 = O.ev_accumulate(ev, 240, 180, 1.0, mode=1, normalize=True)
    ref = py_gauss_splat(zip(ev["x"].tolist(), ev["y"].tolist()), 240, 180, 1.0)
    err = float(np.abs(ref - img).max()) / float(img.max())
    assert err < 2e-6, err           # only expf rounding (python double exp -> f32 vs glibc expf) may differ
    assert mn == 0.0 and abs(mx - img.max()) == 0.0
    # SE3 with identity pose must equal the plain splat up to the double<->float round trip of (u,v)
    K = np.array([199.09, 198.83, 132.19, 110.71], np.float32)
    T = np.eye(4, dtype=np.float32)
    img3, _, _ = O.ev_accumulate(ev, 240, 180, 1.0, mode=2, Tcw=T, depth=1.0, K=K)
    assert float(np.abs(img3 - img).max()) / float(img.max()) < 1e-3
    # nearest
    imgn, (mn0, mx0), u8n = O.ev_accumulate(ev, 240, 180, 1.0, mode=0, normalize=True)
    cnt = np.zeros((180, 240), np.int64)
    px = np.floor(ev["x"].astype(np.float64) + 0.5).astype(int); py = np.floor(ev["y"].astype(np.float64) + 0.5).astype(int)
    px = np.where(ev["x"] < 0, -np.floor(-ev["x"].astype(np.float64) + 0.5).astype(int), px)
    py = np.where(ev["y"] < 0, -np.floor(-ev["y"].astype(np.float64) + 0.5).astype(int), py)
    ok = (px >= 0) & (px < 240) & (py >= 0) & (py < 180)
    np.add.at(cnt, (py[ok], px[ok]), 1)
    assert np.abs(imgn - cnt * 0.001).max() < 1e-6
    Trot = synth.rotation_tcw([0.02, -0.03, 0.05])
    img4, _, _ = O.ev_accumulate(ev, 240, 180, 1.0, mode=2, Tcw=Trot, depth=1.0, K=K)
    img5, _, _ = O.ev_accumulate(ev, 240, 180, 1.0, mode=3, K=K, se2=np.array([0.03, 0.01, -0.02], np.float32))
    np.savez_compressed(os.path.join(HERE, "events_2000.npz"), seed=np.array([11]), gauss=img, gauss_u8=u8, py_gauss=ref,
                        nearest=imgn, K=K, Trot=Trot, se3=img4, se2=img5, se2_params=np.array([0.03, 0.01, -0.02], np.float32))
    print("events pinned: gauss vs python-f32 rel err %.2e, peak %.4f" % (err, float(img.max())))


if __name__ == "__main__":
    pin_primitives()
    pin_pipeline()
    pin_events()
    print("golden fixtures written to", HERE)
