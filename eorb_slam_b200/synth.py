"""Deterministic synthetic inputs of the shapes BASELINE.json names.

There is no network for datasets, so every test and bench input is generated here from a seed:
  * frames    — EuRoC-shaped 8-bit images: background 128 + ~400 random axis-aligned rectangles
                (side 6..39 px, grey 0..255) + uniform noise -6..+6, clamped; gives 7-8x the per-level
                quota of FAST candidates on every pyramid level and a few dozen minThFAST fallback cells.
  * events    — DAVIS-shaped streams: positions on moving straight edges + 10 % uniform noise,
                sub-pixel float coordinates, strictly increasing timestamps, Bernoulli(0.5) polarity.
                Record layout = EventData {double ts; float x; float y; bool p} (24 B,
                reference include/Event/EventData.h:36-58).
  * descriptors — uniform random 256-bit rows plus planted near-duplicates for the Hamming search.
"""
from __future__ import annotations

import numpy as np

EVENT_DTYPE = np.dtype(
    {"names": ["ts", "x", "y", "p"], "formats": ["<f8", "<f4", "<f4", "u1"], "offsets": [0, 8, 12, 16], "itemsize": 24}
)
KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")]
)
MATCH_DTYPE = np.dtype([("best_dist", "<i4"), ("best_idx", "<i4"), ("second_dist", "<i4"), ("accepted", "<i4")])


def make_frame(seed: int, w: int = 752, h: int = 480, nrect: int = 400, noise: int = 6, kind: str = "textured") -> np.ndarray:
    """One synthetic u8 frame (h, w)."""
    rng = np.random.default_rng(np.random.SeedSequence([0xE0B5, int(seed)]))
    if kind == "zero":
        return np.zeros((h, w), np.uint8)
    if kind == "flat":  # low texture: exercises the minThFAST fallback and empty cells
        img = np.full((h, w), 110, np.int16)
        img += rng.integers(-2, 3, size=(h, w), dtype=np.int16)
        # a handful of faint blobs so that some cells only fire at the fallback threshold
        for _ in range(30):
            x0 = int(rng.integers(0, w - 12)); y0 = int(rng.integers(0, h - 12))
            img[y0:y0 + int(rng.integers(4, 12)), x0:x0 + int(rng.integers(4, 12))] += int(rng.integers(9, 16))
        return np.clip(img, 0, 255).astype(np.uint8)
    img = np.full((h, w), 128, np.int16)
    nrect = max(1, int(nrect * (w * h) / (752 * 480)))
    xs = rng.integers(0, w, nrect); ys = rng.integers(0, h, nrect)
    ws = rng.integers(6, 40, nrect); hs = rng.integers(6, 40, nrect)
    gs = rng.integers(0, 256, nrect)
    for x0, y0, rw, rh, g in zip(xs, ys, ws, hs, gs):
        img[y0:y0 + rh, x0:x0 + rw] = g
    img += rng.integers(-noise, noise + 1, size=(h, w), dtype=np.int16)
    return np.clip(img, 0, 255).astype(np.uint8)


def make_frames(n: int, seed0: int = 0, w: int = 752, h: int = 480, unique: int | None = None) -> np.ndarray:
    """(n, h, w) u8 batch; if `unique` < n the first `unique` frames are generated and cycled with a
    per-frame circular shift so that every frame is still distinct."""
    unique = n if unique is None else min(unique, n)
    base = np.stack([make_frame(seed0 + i, w, h) for i in range(unique)])
    if unique == n:
        return base
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        out[i] = np.roll(base[i % unique], shift=(i // unique) * 7, axis=1)
    return out


def make_events(n: int, seed: int, w: int = 240, h: int = 180, n_edges: int = 30, noise_frac: float = 0.1,
                mean_dt: float = 1e-6, t0: float = 0.0) -> np.ndarray:
    """n events as a structured array with EVENT_DTYPE (24-byte records)."""
    rng = np.random.default_rng(np.random.SeedSequence([0xE7E7, int(seed)]))
    ev = np.zeros(n, EVENT_DTYPE)
    ts = t0 + np.cumsum(rng.exponential(mean_dt, n) + 1e-9)
    ev["ts"] = ts
    # moving straight edges: point p0 + s*dir, drifting with velocity vel over time
    p0 = rng.uniform([0, 0], [w, h], size=(n_edges, 2))
    ang = rng.uniform(0, np.pi, n_edges)
    length = rng.uniform(20, 0.6 * min(w, h), n_edges)
    drift = rng.uniform(-6.0, 6.0, size=(n_edges, 2))   # total edge motion (px) over the stream
    which = rng.integers(0, n_edges, n)
    s = rng.uniform(-0.5, 0.5, n) * length[which]
    frac = (ts - t0) / max(ts[-1] - t0, 1e-12)
    x = p0[which, 0] + s * np.cos(ang[which]) + drift[which, 0] * frac + rng.normal(0, 0.35, n)
    y = p0[which, 1] + s * np.sin(ang[which]) + drift[which, 1] * frac + rng.normal(0, 0.35, n)
    noise = rng.random(n) < noise_frac
    x[noise] = rng.uniform(0, w, int(noise.sum()))
    y[noise] = rng.uniform(0, h, int(noise.sum()))
    # keep a few events outside the image so the in-image test is exercised
    ev["x"] = np.clip(x, -4.0, w + 3.0).astype(np.float32)
    ev["y"] = np.clip(y, -4.0, h + 3.0).astype(np.float32)
    ev["p"] = (rng.random(n) < 0.5).astype(np.uint8)
    return ev


def rotation_tcw(omega_xyz, translation=(0.0, 0.0, 0.0)) -> np.ndarray:
    """4x4 float32 Tcw from a rotation vector (Rodrigues) and a translation."""
    w = np.asarray(omega_xyz, np.float64)
    th = float(np.linalg.norm(w))
    K = np.zeros((3, 3))
    if th > 0:
        k = w / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = translation
    return T.astype(np.float32)


def make_descriptor_db(n: int, seed: int) -> np.ndarray:
    """(n, 32) u8 uniform random descriptors."""
    rng = np.random.default_rng(np.random.SeedSequence([0xDB, int(seed)]))
    return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)


def make_queries(db: np.ndarray, nq: int, seed: int, max_flips: int = 40, plant_duplicates: bool = True):
    """nq query rows: first half = DB rows with k~U{0..max_flips} bit flips (true matches), second half
    uniform random (best ~ 90-100 -> rejected by TH_LOW).  Returns (queries, src_idx) with src_idx=-1 for
    the random half.  With plant_duplicates the k=0 queries' source rows are also copied to a second DB
    index (done by the caller through `plant_duplicate_rows`) so ties on distance are exercised."""
    rng = np.random.default_rng(np.random.SeedSequence([0x9E, int(seed)]))
    n = db.shape[0]
    nm = nq // 2
    src = rng.integers(0, n, nm)
    q = np.empty((nq, 32), np.uint8)
    bits = np.unpackbits(db[src], axis=1)
    flips = rng.integers(0, max_flips + 1, nm)
    for i in range(nm):
        if flips[i]:
            pos = rng.choice(256, int(flips[i]), replace=False)
            bits[i, pos] ^= 1
    q[:nm] = np.packbits(bits, axis=1)
    q[nm:] = rng.integers(0, 256, size=(nq - nm, 32), dtype=np.uint8)
    src_idx = np.full(nq, -1, np.int64)
    src_idx[:nm] = src
    return q, src_idx


def plant_duplicate_rows(db: np.ndarray, src_rows: np.ndarray, seed: int) -> np.ndarray:
    """Copy each of `src_rows` to another random index (in place); returns the destination indices."""
    rng = np.random.default_rng(np.random.SeedSequence([0xD0, int(seed)]))
    dst = rng.integers(0, db.shape[0], len(src_rows))
    db[dst] = db[src_rows]
    return dst


def make_keypoint_frame_pair(n1: int, n2: int, seed: int, w: int = 752, h: int = 480, nlevels: int = 8, shift=(7.0, -4.0),
                             max_flips: int = 48, twin_frac: float = 0.15):
    """Two synthetic frames' keypoints + descriptors for the guided-matching tests (SearchForInitialization):
    frame 2 = frame 1 moved by `shift` + jitter with up to max_flips flipped descriptor bits, plus unrelated keypoints;
    a fraction of frame-1 keypoints get a near-twin (close position, nearly the same descriptor) so that two queries
    compete for the same frame-2 keypoint (the vnMatches21 take-over path), some coordinates fall outside the image
    bounds (PosInGrid rejects them) and ~half the keypoints sit on levels > 0 (skipped / filtered).
    Returns (kps1, desc1, kps2, desc2, bounds4)."""
    rng = np.random.default_rng(seed)
    def kp_array(n):
        k = np.zeros(n, KEYPOINT_DTYPE)
        k["x"] = rng.uniform(-3.0, w + 3.0, n).astype(np.float32); k["y"] = rng.uniform(-3.0, h + 3.0, n).astype(np.float32)
        k["octave"] = np.where(rng.random(n) < 0.55, 0, rng.integers(1, nlevels, n)).astype(np.int32)
        k["size"] = 31.0; k["angle"] = rng.uniform(0, 360, n).astype(np.float32); k["response"] = rng.integers(8, 120, n)
        k["class_id"] = -1
        return k
    kps1 = kp_array(n1)
    desc1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    ntw = int(n1 * twin_frac)
    if ntw > 0:   # twins: the second half of the twin pairs copies the first half with tiny changes
        src = rng.choice(n1 // 2, ntw, replace=False); dst = n1 // 2 + rng.choice(n1 - n1 // 2, ntw, replace=False)
        kps1["x"][dst] = kps1["x"][src] + rng.uniform(-2, 2, ntw).astype(np.float32)
        kps1["y"][dst] = kps1["y"][src] + rng.uniform(-2, 2, ntw).astype(np.float32)
        kps1["octave"][dst] = kps1["octave"][src]
        desc1[dst] = desc1[src]
        for d in dst[: ntw // 2]:   # half of the twins differ in a few bits, the others are exact duplicates (distance ties)
            b = rng.integers(0, 256, 3); desc1[d, b // 8] ^= (1 << (b % 8)).astype(np.uint8)
    kps2 = kp_array(n2)
    desc2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    nm = min(n1, n2) * 3 // 4
    a = rng.choice(n1, nm, replace=False); b = rng.choice(n2, nm, replace=False)
    kps2["x"][b] = kps1["x"][a] + np.float32(shift[0]) + rng.normal(0, 2.0, nm).astype(np.float32)
    kps2["y"][b] = kps1["y"][a] + np.float32(shift[1]) + rng.normal(0, 2.0, nm).astype(np.float32)
    kps2["octave"][b] = kps1["octave"][a]
    kps2["angle"][b] = np.mod(kps1["angle"][a] + rng.normal(3.0, 4.0, nm).astype(np.float32), np.float32(360.0)).astype(np.float32)
    desc2[b] = desc1[a]
    for i, row in enumerate(b):
        k = int(rng.integers(0, max_flips + 1))
        bits = rng.choice(256, k, replace=False)
        np.bitwise_xor.at(desc2[row], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    bounds = np.array([0.0, 0.0, float(w), float(h)], np.float32)
    return kps1, desc1, kps2, desc2, bounds


def make_vocabulary(k: int = 10, L: int = 4, seed: int = 0, scoring: int = 0, weighting: int = 0, prune: float = 0.08, zero_weight: float = 0.02):
    """Synthetic DBoW2 vocabulary tree in the flat form TemplatedVocabulary::loadFromTextFile builds
    (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1330-1417; the real ORBvoc.txt — k = 10, L = 6 — is a missing blob):
    node 0 is the root, a node's children are created together (contiguous ids, parent id < child id), a child descriptor is
    its parent's with random bits flipped (so the descent is meaningful), a fraction `prune` of the inner nodes stays a leaf
    early or gets fewer than k children, some siblings share a descriptor (distance ties: the first child wins) and a
    fraction `zero_weight` of the words has weight 0 (stopped words are skipped by transform)."""
    rng = np.random.default_rng(seed)
    parent = [0]; desc = [np.zeros(32, np.uint8)]; leaf = [0]; level = [0]
    frontier = [0]
    for lv in range(1, L + 1):
        nxt = []
        for p in frontier:
            nch = k if rng.random() > prune else int(rng.integers(1, k + 1))
            base = desc[p] if lv > 1 else None
            for c in range(nch):
                if base is None:
                    d = rng.integers(0, 256, 32, dtype=np.uint8)
                else:
                    d = base.copy()
                    bits = rng.choice(256, int(rng.integers(8, 64 // lv + 9)), replace=False)
                    np.bitwise_xor.at(d, bits // 8, (1 << (bits % 8)).astype(np.uint8))
                if c > 0 and rng.random() < 0.03:
                    d = desc[-1].copy()                       # twin sibling
                nid = len(parent)
                parent.append(p); desc.append(d); level.append(lv)
                is_leaf = lv == L or (lv > 1 and rng.random() < prune)
                leaf.append(1 if is_leaf else 0)
                if not is_leaf:
                    nxt.append(nid)
        frontier = nxt
    n = len(parent)
    weight = np.zeros(n, np.float64)
    lf = np.array(leaf, bool)
    weight[lf] = np.log(rng.uniform(2.0, 4000.0, int(lf.sum())))          # idf-like
    weight[lf & (rng.random(n) < zero_weight)] = 0.0
    return dict(k=k, L=L, scoring=scoring, weighting=weighting, parent=np.array(parent, np.int32), is_leaf=np.array(leaf, np.uint8),
                desc=np.stack(desc).astype(np.uint8), weight=weight, level=np.array(level, np.int32))


def make_vocabulary_features(voc, n: int, seed: int, max_flips: int = 30) -> np.ndarray:
    """n descriptors: leaves of the vocabulary with up to max_flips flipped bits (several per word) + 10 % random rows"""
    rng = np.random.default_rng(seed)
    leaves = np.flatnonzero(voc["is_leaf"])
    src = rng.choice(leaves, n)
    src[: n // 5] = src[n // 5: 2 * (n // 5)]                 # repeated words -> accumulated weights
    f = voc["desc"][src].copy()
    for i in range(n):
        bits = rng.choice(256, int(rng.integers(0, max_flips + 1)), replace=False)
        np.bitwise_xor.at(f[i], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    rnd = rng.random(n) < 0.1
    f[rnd] = rng.integers(0, 256, (int(rnd.sum()), 32), dtype=np.uint8)
    return f


def make_vocabulary_regular(k: int = 10, L: int = 6, seed: int = 0):
    """A complete k-ary tree of depth L in level order (vectorised; k = 10, L = 6 is the size of ORBvoc.txt: 1 111 111 nodes,
    10^6 words, 35 MB of descriptors).  Child descriptor = parent's with ~1/8 of ... bits flipped less at deeper levels."""
    rng = np.random.default_rng(seed)
    parents = [np.zeros(1, np.int32)]; descs = [np.zeros((1, 32), np.uint8)]
    first = 0; count = 1
    for lv in range(1, L + 1):
        n = count * k
        par = first + np.repeat(np.arange(count, dtype=np.int32), k)
        if lv == 1:
            d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        else:
            flip = rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
            for _ in range(lv - 1):
                flip &= rng.integers(0, 256, (n, 32), dtype=np.uint8)
            d = np.repeat(descs[-1], k, axis=0) ^ flip
        parents.append(par); descs.append(d)
        first += count; count = n
    parent = np.concatenate(parents); desc = np.concatenate(descs)
    nn = len(parent)
    leaf = np.zeros(nn, np.uint8); leaf[nn - count:] = 1
    weight = np.zeros(nn, np.float64); weight[nn - count:] = np.log(rng.uniform(2.0, 4000.0, count))
    return dict(k=k, L=L, scoring=0, weighting=0, parent=parent, is_leaf=leaf, desc=desc, weight=weight)


def make_projection_case(n1: int, n2: int, seed: int, w: int = 752, h: int = 480, nlevels: int = 8, K=(458.654, 457.296, 367.215, 248.375),
                         shift=(7.0, -4.0), valid_frac: float = 0.7, zero_obs_frac: float = 0.05):
    """Inputs of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono=True) for the guided-matching tests: a
    keypoint frame pair (make_keypoint_frame_pair) plus, per last-frame keypoint, a camera-frame point x3Dc that projects onto the
    keypoint moved by `shift` (so the search window contains its counterpart), a validity flag (map point present, not an
    outlier), an observation count (a few are 0: such claims can be overwritten) and the map point's descriptor.  A few points
    lie behind the camera or project outside the image.  Returns a dict."""
    rng = np.random.default_rng(seed + 1000)
    k1, d1, k2, d2, bounds = make_keypoint_frame_pair(n1, n2, seed, w, h, nlevels, shift)
    fx, fy, cx, cy = (np.float32(v) for v in K)
    z = rng.uniform(1.0, 10.0, n1).astype(np.float32)
    z[rng.random(n1) < 0.03] *= np.float32(-1.0)                       # behind the camera
    u = k1["x"] + np.float32(shift[0]) + rng.normal(0, 1.5, n1).astype(np.float32)
    v = k1["y"] + np.float32(shift[1]) + rng.normal(0, 1.5, n1).astype(np.float32)
    far = rng.random(n1) < 0.03
    u[far] += np.float32(2000.0)                                       # projects outside the image
    x3 = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1).astype(np.float32)
    valid = (rng.random(n1) < valid_frac).astype(np.uint8)
    obs = rng.integers(1, 9, n1).astype(np.int32)
    obs[rng.random(n1) < zero_obs_frac] = 0
    sf = np.float32(1.2) ** np.arange(nlevels, dtype=np.float32)
    scale = np.ones(nlevels, np.float32)
    for i in range(1, nlevels):
        scale[i] = np.float32(np.float64(scale[i - 1]) * np.float64(np.float32(1.2)))
    return dict(x3Dc=x3, valid1=valid, obs1=obs, kps1=k1, descMP=d1, kps2=k2, desc2=d2, bounds=bounds, K=np.array(K, np.float32),
                scale_factors=scale)


TRACK_POINT_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("view_cos", "<f4"), ("depth", "<f4"), ("scale_level", "<i4"),
                              ("observations", "<i4"), ("in_view", "u1"), ("bad", "u1"), ("pad", "u1", (2,))])


def make_local_map_case(n1: int, n2: int, seed: int, w: int = 752, h: int = 480, nlevels: int = 8, shift=(3.0, -2.0),
                        in_view_frac: float = 0.8, zero_obs_frac: float = 0.05, held_frac: float = 0.2):
    """Inputs of ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) (Tracking::SearchLocalPoints): n1 local
    map points whose tracking fields (what Frame::isInFrustum fills) put them near the keypoints of a frame of n2 keypoints
    (make_keypoint_frame_pair: frame 1 plays the map points, its descriptors are theirs), a predicted level near the keypoint's
    octave, viewing cosines on both sides of 0.998 (incl. exactly float(0.998)), some points out of view / bad / far / without
    observations, and a fraction of the frame's slots already holding a point with observations."""
    rng = np.random.default_rng(seed + 2000)
    k1, d1, k2, d2, bounds = make_keypoint_frame_pair(n1, n2, seed, w, h, nlevels, shift)
    pts = np.zeros(n1, TRACK_POINT_DTYPE)
    pts["proj_x"] = k1["x"] + np.float32(shift[0]) + rng.normal(0, 1.0, n1).astype(np.float32)
    pts["proj_y"] = k1["y"] + np.float32(shift[1]) + rng.normal(0, 1.0, n1).astype(np.float32)
    vc = rng.uniform(0.99, 1.0, n1).astype(np.float32)
    vc[rng.random(n1) < 0.05] = np.float32(0.998)
    pts["view_cos"] = vc
    pts["depth"] = rng.uniform(0.5, 30.0, n1).astype(np.float32)
    pts["scale_level"] = np.clip(k1["octave"] + rng.integers(-1, 2, n1), -1, nlevels).astype(np.int32)
    obs = rng.integers(1, 9, n1).astype(np.int32)
    obs[rng.random(n1) < zero_obs_frac] = 0
    pts["observations"] = obs
    pts["in_view"] = (rng.random(n1) < in_view_frac).astype(np.uint8)
    pts["bad"] = (rng.random(n1) < 0.03).astype(np.uint8)
    held = (rng.random(n2) < held_frac).astype(np.uint8)
    scale = np.ones(nlevels, np.float32)
    for i in range(1, nlevels):
        scale[i] = np.float32(np.float64(scale[i - 1]) * np.float64(np.float32(1.2)))
    return dict(pts=pts, descMP=d1, kps2=k2, desc2=d2, held2=held, bounds=bounds, scale_factors=scale)
