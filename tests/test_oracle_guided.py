"""Pins oracle/guided_oracle.cc (Frame grid, GetFeaturesInArea, ORBmatcher::SearchForInitialization) against an
independent pure-Python restatement written from the same reference lines (src/Frame.cc:431-460, 709-793;
src/ORBmatcher.cc:714-831, 2314-2355).  The reference ships no tests for these functions."""
import math

import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth

F = np.float32
COLS, ROWS = 64, 48


def _round_away(v):
    v = float(v)
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def py_grid(kps, bounds):
    minx, miny, maxx, maxy = (F(b) for b in bounds)
    winv = F(COLS) / (maxx - minx); hinv = F(ROWS) / (maxy - miny)
    grid = [[[] for _ in range(ROWS)] for _ in range(COLS)]
    for i in range(len(kps)):
        px = _round_away((F(kps["x"][i]) - minx) * winv); py = _round_away((F(kps["y"][i]) - miny) * hinv)
        if px < 0 or px >= COLS or py < 0 or py >= ROWS:
            continue
        grid[px][py].append(i)
    return grid, (minx, miny, winv, hinv)


def py_area(kps, grid, geom, x, y, r, min_level, max_level):
    minx, miny, winv, hinv = geom
    x, y, r = F(x), F(y), F(r)
    out = []
    c0 = max(0, int(math.floor((x - minx - r) * winv)))
    if c0 >= COLS:
        return out
    c1 = min(COLS - 1, int(math.ceil((x - minx + r) * winv)))
    if c1 < 0:
        return out
    r0 = max(0, int(math.floor((y - miny - r) * hinv)))
    if r0 >= ROWS:
        return out
    r1 = min(ROWS - 1, int(math.ceil((y - miny + r) * hinv)))
    if r1 < 0:
        return out
    check = (min_level > 0) or (max_level >= 0)
    for ix in range(c0, c1 + 1):
        for iy in range(r0, r1 + 1):
            for j in grid[ix][iy]:
                if check:
                    lv = int(kps["octave"][j])
                    if lv < min_level or (max_level >= 0 and lv > max_level):
                        continue
                if abs(F(kps["x"][j]) - x) < r and abs(F(kps["y"][j]) - y) < r:
                    out.append(j)
    return out


def _ham(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def py_search_init(kps1, d1, kps2, d2, bounds, prev, window, nnratio, check_ori):
    INT_MAX = 2 ** 31 - 1
    grid, geom = py_grid(kps2, bounds)
    n1, n2 = len(kps1), len(kps2)
    m12 = [-1] * n1; m21 = [-1] * n2; md = [INT_MAX] * n2
    hist = [[] for _ in range(30)]
    prev = np.array(prev, np.float32, copy=True)
    nm = 0
    for i1 in range(n1):
        if kps1["octave"][i1] > 0:
            continue
        cand = py_area(kps2, grid, geom, prev[i1, 0], prev[i1, 1], float(window), 0, 0)
        if not cand:
            continue
        best, best2, bidx = INT_MAX, INT_MAX, -1
        for i2 in cand:
            dist = _ham(d1[i1], d2[i2])
            if md[i2] <= dist:
                continue
            if dist < best:
                best2, best, bidx = best, dist, i2
            elif dist < best2:
                best2 = dist
        if best <= 50 and F(best) < F(best2) * F(nnratio):
            if m21[bidx] >= 0:
                m12[m21[bidx]] = -1; nm -= 1
            m12[i1] = bidx; m21[bidx] = i1; md[bidx] = best; nm += 1
            if check_ori:
                rot = F(kps1["angle"][i1]) - F(kps2["angle"][bidx])
                if rot < 0:
                    rot = rot + F(360.0)
                b = _round_away(F(rot) * (F(1.0) / F(30)))
                hist[0 if b == 30 else b].append(i1)
    if check_ori:
        mx = [0, 0, 0]; ind = [-1, -1, -1]
        for i in range(30):
            s = len(hist[i])
            if s > mx[0]:
                mx = [s, mx[0], mx[1]]; ind = [i, ind[0], ind[1]]
            elif s > mx[1]:
                mx = [mx[0], s, mx[1]]; ind = [ind[0], i, ind[1]]
            elif s > mx[2]:
                mx[2] = s; ind[2] = i
        if mx[1] < F(0.1) * F(mx[0]):
            ind[1] = ind[2] = -1
        elif mx[2] < F(0.1) * F(mx[0]):
            ind[2] = -1
        for i in range(30):
            if i in ind:
                continue
            for i1 in hist[i]:
                if m12[i1] >= 0:
                    m12[i1] = -1; nm -= 1
    for i1 in range(n1):
        if m12[i1] >= 0:
            prev[i1] = (kps2["x"][m12[i1]], kps2["y"][m12[i1]])
    return nm, np.array(m12, np.int32), prev


@pytest.mark.parametrize("seed,bounds", [(1, None), (2, (-12.5, -9.25, 771.0, 493.5))])
def test_grid_and_features_in_area_match_python(seed, bounds):
    k1, _, k2, _, b = synth.make_keypoint_frame_pair(600, 700, seed)
    if bounds is not None:
        b = np.array(bounds, np.float32)   # undistorted image bounds need not start at 0 (Frame.cc:855-860)
    cs, ci = O.frame_grid(k2, b)
    grid, geom = py_grid(k2, b)
    flat = [i for c in range(COLS) for r in range(ROWS) for i in grid[c][r]]
    assert ci.tolist() == flat
    assert cs[-1] == len(flat) and all(cs[c * ROWS + r + 1] - cs[c * ROWS + r] == len(grid[c][r]) for c in range(COLS) for r in range(ROWS))
    rng = np.random.default_rng(seed)
    for q in range(150):
        x, y = rng.uniform(-60, 820), rng.uniform(-60, 540)
        r = float(rng.choice([5.0, 15.0, 100.0, 900.0]))
        lo, hi = [(0, -1), (0, 0), (2, 4), (1, -1), (3, 2)][q % 5]
        assert O.features_in_area(k2, b, cs, ci, x, y, r, lo, hi).tolist() == py_area(k2, grid, geom, x, y, r, lo, hi)


@pytest.mark.parametrize("seed,window,ratio,ori", [(3, 100, 0.9, True), (4, 30, 0.7, True), (5, 100, 0.9, False), (6, 15, 0.95, True)])
def test_search_for_initialization_matches_python(seed, window, ratio, ori):
    k1, d1, k2, d2, b = synth.make_keypoint_frame_pair(500, 520, seed)
    prev = np.stack([k1["x"], k1["y"]], 1)   # Tracking.cc: vbPrevMatched starts as the frame-1 keypoint positions
    n, m12, p = O.search_for_initialization(k1, d1, k2, d2, b, prev, window, ratio, ori)
    en, em12, ep = py_search_init(k1, d1, k2, d2, b, prev, window, ratio, ori)
    assert n == en and np.array_equal(m12, em12) and p.tobytes() == ep.tobytes()
    assert n > 50                                     # the generator really produces matches ...
    if window >= 100:
        # ... and exercises the take-over path: some frame-2 keypoint was claimed by more than one query
        assert n == int((m12 >= 0).sum())


def test_search_for_initialization_second_round_uses_updated_prev_matched():
    """the reference calls the search again with the updated vbPrevMatched when initialisation fails"""
    k1, d1, k2, d2, b = synth.make_keypoint_frame_pair(400, 400, 9)
    prev = np.stack([k1["x"], k1["y"]], 1)
    n, m12, p = O.search_for_initialization(k1, d1, k2, d2, b, prev, 100, 0.9, True)
    n2, m12b, p2 = O.search_for_initialization(k1, d1, k2, d2, b, p, 100, 0.9, True)
    en2, em12b, ep2 = py_search_init(k1, d1, k2, d2, b, p, 100, 0.9, True)
    assert n2 == en2 and np.array_equal(m12b, em12b) and p2.tobytes() == ep2.tobytes()


def py_search_by_projection(c, th, check_ori):
    k1, k2 = c["kps1"], c["kps2"]
    grid, geom = py_grid(k2, c["bounds"])
    fx, fy, cx, cy = (F(v) for v in c["K"])
    b = [F(v) for v in c["bounds"]]
    n2 = len(k2)
    mc = [-1] * n2
    hist = [[] for _ in range(30)]
    nm = 0
    nl = len(c["scale_factors"])
    for i in range(len(k1)):
        if not c["valid1"][i]:
            continue
        xc, yc, zc = (F(v) for v in c["x3Dc"][i])
        if zc == 0 or F(1.0 / float(zc)) < 0:
            continue
        u = fx * xc / zc + cx; v = fy * yc / zc + cy
        if u < b[0] or u > b[2] or v < b[1] or v > b[3]:
            continue
        octv = int(k1["octave"][i])
        radius = F(th) * F(c["scale_factors"][min(max(octv, 0), nl - 1)])
        cand = py_area(k2, grid, geom, u, v, radius, octv - 1, octv + 1)
        best, bidx = 256, -1
        for i2 in cand:
            if mc[i2] >= 0 and c["obs1"][mc[i2]] > 0:
                continue
            d = _ham(c["descMP"][i], c["desc2"][i2])
            if d < best:
                best, bidx = d, i2
        if best <= 100:
            mc[bidx] = i; nm += 1
            if check_ori:
                rot = F(k1["angle"][i]) - F(k2["angle"][bidx])
                if rot < 0:
                    rot = rot + F(360.0)
                bb = _round_away(F(rot) * (F(1.0) / F(30)))
                hist[0 if bb == 30 else bb].append(bidx)
    if check_ori:
        mx = [0, 0, 0]; ind = [-1, -1, -1]
        for i in range(30):
            s = len(hist[i])
            if s > mx[0]:
                mx = [s, mx[0], mx[1]]; ind = [i, ind[0], ind[1]]
            elif s > mx[1]:
                mx = [mx[0], s, mx[1]]; ind = [ind[0], i, ind[1]]
            elif s > mx[2]:
                mx[2] = s; ind[2] = i
        if mx[1] < F(0.1) * F(mx[0]):
            ind[1] = ind[2] = -1
        elif mx[2] < F(0.1) * F(mx[0]):
            ind[2] = -1
        for i in range(30):
            if i not in ind:
                for idx in hist[i]:
                    mc[idx] = -1; nm -= 1
    return nm, np.array(mc, np.int32)


@pytest.mark.parametrize("seed,th,ori,zero_obs", [(21, 15.0, True, 0.05), (22, 7.0, True, 0.0), (23, 15.0, False, 0.3), (24, 40.0, True, 0.5)])
def test_search_by_projection_matches_python(seed, th, ori, zero_obs):
    c = synth.make_projection_case(500, 520, seed, zero_obs_frac=zero_obs)
    n, mc = O.search_by_projection(c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], c["K"],
                                   c["scale_factors"], th, ori)
    en, emc = py_search_by_projection(c, th, ori)
    assert n == en and np.array_equal(mc, emc)
    assert (mc >= 0).sum() > 50
    if zero_obs == 0.0 and not ori:
        assert n == int((mc >= 0).sum())


def py_search_local_points(c, th, far_points, th_far, nnratio):
    """independent restatement of ORBmatcher.cc:44-148 (monocular) with Python lists"""
    pts, k2 = c["pts"], c["kps2"]
    grid, geom = py_grid(k2, c["bounds"])
    n2 = len(k2)
    holder = [(-2 if c["held2"][i] else -1) for i in range(n2)]   # -2: holds a point with observations on entry
    nl = len(c["scale_factors"])
    nm = 0
    for i in range(len(pts)):
        p = pts[i]
        if not p["in_view"] or (far_points and F(p["depth"]) > F(th_far)) or p["bad"]:
            continue
        lvl = int(p["scale_level"])
        r = F(2.5) if float(p["view_cos"]) > 0.998 else F(4.0)
        if float(F(th)) != 1.0:
            r = r * F(th)
        cand = py_area(k2, grid, geom, F(p["proj_x"]), F(p["proj_y"]), F(r * F(c["scale_factors"][min(max(lvl, 0), nl - 1)])), lvl - 1, lvl)
        seen = []
        for i2 in cand:
            h = holder[i2]
            if h == -2 or (h >= 0 and pts["observations"][h] > 0):
                continue
            seen.append((_ham(c["descMP"][i], c["desc2"][i2]), len(seen), i2))
        if not seen:
            continue
        seen.sort()                                             # first two of the (distance, visiting position) order
        bd, _, bi = seen[0]
        bl = int(k2["octave"][bi])
        bd2, bl2 = (seen[1][0], int(k2["octave"][seen[1][2]])) if len(seen) > 1 else (256, -1)
        if bd <= 100 and not (bl == bl2 and F(bd) > F(nnratio) * F(bd2)):
            holder[bi] = i; nm += 1
    return nm, np.array([h if h >= 0 else -1 for h in holder], np.int32)


@pytest.mark.parametrize("seed,th,far,ratio,zero_obs", [(31, 1.0, False, 0.8, 0.05), (32, 3.0, True, 0.8, 0.0), (33, 5.0, False, 0.6, 0.4),
                                                       (34, 15.0, True, 0.9, 0.2)])
def test_search_by_projection_map_points_matches_python(seed, th, far, ratio, zero_obs):
    c = synth.make_local_map_case(600, 520, seed, zero_obs_frac=zero_obs)
    n, mc = O.search_by_projection_map_points(c["pts"], c["descMP"], c["kps2"], c["desc2"], c["held2"], c["bounds"], c["scale_factors"], th, far,
                                              20.0, ratio)
    en, emc = py_search_local_points(c, th, far, 20.0, ratio)
    assert n == en and np.array_equal(mc, emc)
    assert (mc >= 0).sum() > 40
    assert not np.any((mc >= 0) & (c["held2"] != 0))           # a held slot is never taken
    n0, mc0 = O.search_by_projection_map_points(c["pts"], c["descMP"], c["kps2"], c["desc2"], None, c["bounds"], c["scale_factors"], th, far,
                                                20.0, ratio)
    assert n0 >= n


def py_search_by_bow(k1, d1, valid, fv1, k2, d2, fv2, ratio, ori):
    """independent restatement of ORBmatcher.cc:276-478 (monocular) with dicts keyed by node id, like DBoW2's FeatureVector"""
    m1 = {int(n): [int(x) for x in fv1[2][fv1[1][i]:fv1[1][i + 1]]] for i, n in enumerate(fv1[0])}
    m2 = {int(n): [int(x) for x in fv2[2][fv2[1][i]:fv2[1][i + 1]]] for i, n in enumerate(fv2[0])}
    mf = [-1] * len(k2)
    hist = [[] for _ in range(30)]
    nm = 0
    for node in sorted(set(m1) & set(m2)):
        for ik in m1[node]:
            if not valid[ik]:
                continue
            ds = [(_ham(d1[ik], d2[j]), j) for j in m2[node] if mf[j] < 0]
            if not ds:
                continue
            best = min(ds, key=lambda t: t[0])                 # first of the smallest
            rest = sorted(t[0] for t in ds)
            b2 = rest[1] if len(rest) > 1 else 256
            if best[0] <= 50 and F(best[0]) < F(ratio) * F(b2):
                mf[best[1]] = ik; nm += 1
                if ori:
                    rot = F(k1["angle"][ik]) - F(k2["angle"][best[1]])
                    if rot < 0:
                        rot = rot + F(360.0)
                    bb = _round_away(F(rot) * (F(1.0) / F(30)))
                    hist[0 if bb == 30 else bb].append(best[1])
    if ori:
        mx = [0, 0, 0]; ind = [-1, -1, -1]
        for i in range(30):
            s = len(hist[i])
            if s > mx[0]:
                mx = [s, mx[0], mx[1]]; ind = [i, ind[0], ind[1]]
            elif s > mx[1]:
                mx = [mx[0], s, mx[1]]; ind = [ind[0], i, ind[1]]
            elif s > mx[2]:
                mx[2] = s; ind[2] = i
        if mx[1] < F(0.1) * F(mx[0]):
            ind[1] = ind[2] = -1
        elif mx[2] < F(0.1) * F(mx[0]):
            ind[2] = -1
        for i in range(30):
            if i not in ind:
                for idx in hist[i]:
                    mf[idx] = -1; nm -= 1
    return nm, np.array(mf, np.int32)


def bow_case(n1, n2, seed, k=6, L=3, levelsup=2, valid_frac=0.8, max_flips=24):
    """keyframe / frame keypoints + descriptors, their FeatureVectors through the vocabulary oracle, and the map-point flags"""
    k1, d1, k2, d2, _ = synth.make_keypoint_frame_pair(n1, n2, seed, max_flips=max_flips)
    vo = O.VocabOracle(synth.make_vocabulary(k, L, seed))
    t1, t2 = vo.transform(d1, levelsup), vo.transform(d2, levelsup)
    valid = (np.random.default_rng(seed).random(n1) < valid_frac).astype(np.uint8)
    return k1, d1, valid, (t1["fv_nodes"], t1["fv_start"], t1["fv_feats"]), k2, d2, (t2["fv_nodes"], t2["fv_start"], t2["fv_feats"])


@pytest.mark.parametrize("seed,ratio,ori,levelsup", [(41, 0.7, True, 2), (42, 0.9, True, 1), (43, 0.7, False, 3), (44, 0.6, True, 0)])
def test_search_by_bow_matches_python(seed, ratio, ori, levelsup):
    c = bow_case(500, 520, seed, levelsup=levelsup)
    n, mf = O.search_by_bow(*c, ratio, ori)
    en, emf = py_search_by_bow(*c, ratio, ori)
    assert n == en and np.array_equal(mf, emf)
    assert n == int((mf >= 0).sum())                           # one keyframe feature per matched frame keypoint
    if levelsup >= 2:
        assert n > 50
