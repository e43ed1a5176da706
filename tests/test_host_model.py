"""CPU suite: the product's shared host/device sources (eorb_math.cuh, octree_core.cuh) compiled as plain C++
(one 'thread') and checked against the serial oracle.  This validates the data-parallel FORMULATION of the
octree kernel and the scalar arithmetic of the other kernels without a GPU; races are checked on the GPU box."""
import ctypes as C

import numpy as np

import oracle_lib as O


def _octree_case(hm, rng, W, H, N, xs, ys, sc):
    ref = O.distribute_octtree(xs.astype(np.float32), ys.astype(np.float32), sc.astype(np.float32), 0, W, 0, H, N)
    keys = np.ascontiguousarray((xs.astype(np.uint32) | (ys.astype(np.uint32) << 12) | (sc.astype(np.uint32) << 24)).astype(np.uint32))
    nIni = int(np.round(np.float32(W) / np.float32(H)))
    out = np.zeros(max(N + 3, 4 * nIni) + 8, np.uint32)
    r = hm.hm_octree(keys.ctypes.data if len(keys) else None, len(keys), W, H, N, out.ctypes.data, len(out))
    exp = keys[ref] if len(keys) else np.zeros(0, np.uint32)
    return r == len(ref) and np.array_equal(out[:max(r, 0)], exp)


def test_octree_formulation_matches_serial_oracle(host_model):
    rng = np.random.default_rng(2024)
    ncase = 0
    for trial in range(700):
        W = int(rng.integers(20, 760)); H = int(rng.integers(20, 480))
        if round(float(np.float32(W) / np.float32(H))) < 1 or W <= 6 or H <= 6:
            continue
        mode = trial % 4
        n = int(rng.integers(0, 2500)) if mode != 3 else int(rng.integers(0, 40))
        if mode == 1:   # clustered keys: many degenerate (single non-empty child) splits
            cx = rng.integers(3, W - 3, 5); cy = rng.integers(3, H - 3, 5)
            xs = np.clip((cx[rng.integers(0, 5, n)] + rng.normal(0, 6, n)).astype(int), 3, W - 4)
            ys = np.clip((cy[rng.integers(0, 5, n)] + rng.normal(0, 6, n)).astype(int), 3, H - 4)
        else:
            xs = rng.integers(3, W - 3, n); ys = rng.integers(3, H - 3, n)
        pos = np.unique(np.stack([ys, xs], 1), axis=0)
        rng.shuffle(pos)
        ys = pos[:, 0]; xs = pos[:, 1]
        sc = rng.integers(0, 255 if mode != 2 else 4, len(xs))   # mode 2: many response ties
        N = int(rng.integers(0, 1200)) if trial % 7 else int(rng.integers(0, 12))
        assert _octree_case(host_model, rng, W, H, N, xs, ys, sc), (trial, W, H, N, len(xs))
        ncase += 1
    assert ncase > 500


def test_octree_on_real_candidates(host_model):
    from eorb_slam_b200 import synth
    orc = O.OrbOracle()
    orc.extract(synth.make_frame(0))
    quota = orc.features_per_level()
    for l in range(8):
        xs, ys, sc = orc.candidates(l)
        w, h = orc.level_size(l)
        W = w - 2 * 19 + 6; H = h - 2 * 19 + 6
        assert _octree_case(host_model, None, W, H, int(quota[l]), xs, ys, sc), l


def test_fast_arc_score_matches_oracle(host_model):
    rng = np.random.default_rng(5)
    dx = [0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1]
    dy = [3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3]
    for trial in range(300):
        tile = rng.integers(0, 256, (7, 7), dtype=np.uint8)
        if trial % 3 == 0:
            tile[:] = 100; tile[0:4, :] = rng.integers(0, 256)   # structured: a real corner/edge
        ring = np.array([int(tile[3 + dy[k], 3 + dx[k]]) for k in range(16)], np.int32)
        m = host_model.hm_fast_max_arc_min(int(tile[3, 3]), ring.ctypes.data)
        for t in (0, 7, 20):
            xs, ys, sc = O.fast(tile, t, False)
            assert (len(xs) == 1) == (m > t)
            xs, ys, sc = O.fast(tile, t, True)
            if m > t and m - 1 > 0:
                assert len(sc) == 1 and sc[0] == m - 1


def test_scalar_math_matches_oracle(host_model):
    rng = np.random.default_rng(11)
    for _ in range(3000):
        y = float(rng.integers(-3000000, 3000000)); x = float(rng.integers(-3000000, 3000000))
        a = np.float32(host_model.hm_fast_atan2(C.c_float(y), C.c_float(x))); b = np.float32(O.fast_atan2(y, x))
        assert a.view(np.uint32) == b.view(np.uint32)
    a32 = rng.integers(0, 2 ** 32, (100, 8), dtype=np.uint32); b32 = rng.integers(0, 2 ** 32, (100, 8), dtype=np.uint32)
    for i in range(100):
        assert host_model.hm_hamming(a32[i].ctypes.data, b32[i].ctypes.data) == O.descriptor_distance(a32[i].view(np.uint8), b32[i].view(np.uint8))
    # steered BRIEF offsets: no-FMA float arithmetic + round-half-even
    r = C.c_int(); c = C.c_int()
    for _ in range(2000):
        ang = np.float32(rng.uniform(0, 360)) * np.float32(np.pi / 180.0)
        ca = np.float32(np.cos(np.float64(ang))); sa = np.float32(np.sin(np.float64(ang)))
        px = int(rng.integers(-13, 14)); py = int(rng.integers(-13, 14))
        host_model.hm_brief_offset(px, py, C.c_float(ca), C.c_float(sa), C.byref(r), C.byref(c))
        er = int(np.rint(np.float32(np.float32(px) * sa) + np.float32(np.float32(py) * ca)))
        ec = int(np.rint(np.float32(np.float32(px) * ca) - np.float32(np.float32(py) * sa)))
        assert (r.value, c.value) == (er, ec)
