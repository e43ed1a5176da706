#!/usr/bin/env python3
"""Write the REFERENCE-PINNED golden fixtures tests/golden/ref_*.npz + ref_report.json.

Source of truth here is oracle/_ref/libref.so: the reference's own src/ORBextractor.cc, src/Event/EventConversion.cc and
ORBmatcher::DescriptorDistance compiled UNMODIFIED (recipe: `make -C oracle ref`, stand-in headers in oracle/ref_mock/).
It only exists where /root/reference does (the authoring container), so this script runs there and its outputs are
committed; tests/test_ref_pin.py (CPU) demands oracle == these files byte for byte, tests/test_gpu_ref.py (GPU) demands the
same of the CUDA path.

Cases stay inside the domain where the reference is a function of its inputs: margin >= 19 wherever descriptors are
compared (below that computeOrbDescriptor reads outside the blurred level, ORBextractor.cc:119-124), every level large
enough for one FAST cell (else :803-806 divides by zero), aspect ratio such that DistributeOctTree gets >= 1 root (:563).
The octree's pointer-order tie rule (:703) is fixed by the bump allocator of libref (see oracle/ref_api.cc); the report
counts how many keypoints change under glibc malloc and how many descriptor bytes change in an FMA build.

Usage:  python tests/golden/make_ref_golden.py     (asserts oracle == libref on every case, then rewrites the files)
"""
import ast
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
import ref_lib as R  # noqa: E402
from eorb_slam_b200 import synth  # noqa: E402

GOLDEN_NAMES = ["cfg1_seed0", "cfg1_seed1_stereo", "cfg1_flat", "mvsec_346x260", "ethz_240x180_e9", "ev_single_level"]
N_FUZZ = 200
N_OCTREE = 300
FULL_EVENT_FRAMES = (2, 3, 6, 10)


def kpset(k):
    return set(bytes(r) for r in np.ascontiguousarray(k).view(np.uint8).reshape(len(k), 28)) if len(k) else set()


def sha(a) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def level_geometry(w, h, nlev, sf, edge):
    """per-level sizes as ComputePyramid derives them (ORBextractor.cc:1244-1245) through libref's own constructor"""
    t = R.RefOrb(100, sf, nlev, 20, 7, edge, w, h).tables(w, h)
    return t["edge"], [(int(a), int(b)) for a, b in zip(t["level_w"], t["level_h"])]


def reference_defined(w, h, nlev, sf, edge, want_desc):
    """True when the reference's code has no division by zero / out-of-range index / out-of-bounds read for this geometry"""
    if nlev > 1 and sf <= 1.0:
        return False                      # quota series is 0/0 (ORBextractor.cc:447)
    if edge < 0 and w < 752:
        return False                      # adaptive margin below 19
    E, sizes = level_geometry(w, h, nlev, sf, edge)
    if E < 3 or (want_desc and E < 19):
        return False
    for (lw, lh) in sizes:
        bw, bh = lw - 2 * E + 6, lh - 2 * E + 6
        if bw < 30 or bh < 30:
            return False                  # nCols or nRows would be 0 (:803-806)
        if round(np.float32(bw) / np.float32(bh)) < 1:
            return False                  # nIni = 0 (:563)
    return True


def in_bounds_rows(kps, sizes, scale, reach=19):
    """rows whose 31x31 rotated BRIEF taps stay inside the borderless blurred level (always all rows when margin >= 19)"""
    ok = np.ones(len(kps), bool)
    for i, k in enumerate(kps):
        l = int(k["octave"])
        x = k["x"] / scale[l] if l else k["x"]
        y = k["y"] / scale[l] if l else k["y"]
        lw, lh = sizes[l]
        ok[i] = (round(float(x)) >= reach and round(float(y)) >= reach and round(float(x)) < lw - reach and round(float(y)) < lh - reach)
    return ok


def fuzz_case(rng):
    """one random configuration inside the reference-defined domain -> dict of plain python values"""
    while True:
        w = int(rng.integers(120, 900)); h = int(rng.integers(100, 620))
        nlev = int(rng.integers(1, 10))
        sf = float(np.float32(rng.choice([1.1, 1.2, 1.26, 1.5, 2.0]))) if nlev > 1 else 1.0
        nfeat = int(rng.choice([1, 50, 400, 1000, 2500]))
        ini = int(rng.integers(0, 40)); mn = int(rng.integers(0, ini + 1))
        edge = int(rng.choice([19, 19, 21, 25, 31, -1]))
        lap = [(0, 1000), (0, 0), (100, 300)][int(rng.integers(0, 3))]
        want = bool(rng.integers(0, 4) > 0)
        kind = ["textured", "flat", "textured"][int(rng.integers(0, 3))]
        c = dict(w=w, h=h, nlev=nlev, sf=sf, nfeat=nfeat, ini=ini, mn=mn, edge=edge, lap0=lap[0], lap1=lap[1], want=int(want), kind=kind,
                 seed=int(rng.integers(0, 10**6)), nrect=int(rng.integers(5, 500)), noise=int(rng.integers(0, 12)))
        if reference_defined(w, h, nlev, sf, edge, want):
            return c


def fuzz_frame(c):
    return synth.make_frame(c["seed"], c["w"], c["h"], nrect=c["nrect"], noise=c["noise"], kind=c["kind"])


def octree_case(rng):
    """random candidate sets with many equal-sized nodes: duplicated clusters and a lattice give ties at every depth"""
    w = int(rng.integers(60, 760)); h = int(rng.integers(40, int(min(500, w / 0.6)) + 1))
    n = int(rng.choice([0, 1, 2, 7, 60, 400, 1500, 4000]))
    N = int(rng.choice([0, 1, 5, 60, 217, 1000]))
    style = int(rng.integers(0, 3))
    if style == 0:
        x = rng.integers(0, w, n); y = rng.integers(0, h, n)
    elif style == 1:     # lattice: identical occupancy in sibling nodes
        g = max(2, int(rng.integers(2, 12)))
        x = (rng.integers(0, max(1, w // g), n) * g) % w; y = (rng.integers(0, max(1, h // g), n) * g) % h
    else:                # a few dense blobs
        cx = rng.integers(0, w, 6); cy = rng.integers(0, h, 6); k = rng.integers(0, 6, n)
        x = np.clip(cx[k] + rng.integers(-9, 10, n), 0, w - 1); y = np.clip(cy[k] + rng.integers(-9, 10, n), 0, h - 1)
    # FAST never returns the same pixel twice
    xy = np.unique(np.stack([y, x], 1), axis=0) if n else np.zeros((0, 2), np.int64)
    order = rng.permutation(len(xy))
    xy = xy[np.sort(order)] if rng.integers(0, 2) else xy     # row-major (as FAST emits) most of the time
    resp = rng.integers(1, 60 if rng.integers(0, 2) else 4, len(xy)).astype(np.float32)   # few distinct responses: response ties too
    return dict(w=w, h=h, N=N, x=xy[:, 1].astype(np.float32), y=xy[:, 0].astype(np.float32), resp=resp)


def main():
    assert R.can_build(), "needs the reference tree (set EORB_REFERENCE)"
    R.build("ref")
    report = {"libref": "oracle/_ref/libref.so: ORBextractor.cc + EventConversion.cc + DescriptorDistance compiled unmodified",
              "alloc_pin": "bump arena (address order == creation order)", "cases": {}}

    # ------------------------------------------------------------------ 1. the six golden configurations
    for name in GOLDEN_NAMES:
        g = np.load(os.path.join(HERE, "orb_%s.npz" % name))
        fkw = ast.literal_eval(str(g["frame_kw"])); okw = ast.literal_eval(str(g["orb_kw"]))
        img = synth.make_frame(**fkw)
        lap = tuple(int(v) for v in g["lapping"])
        args = (okw["nfeatures"], okw["scale_factor"], okw["nlevels"], okw["ini_th"], okw["min_th"], okw["edge"], fkw["w"], fkw["h"])
        ref = R.RefOrb(*args)
        orc = O.OrbOracle(*args)
        tb = ref.tables(fkw["w"], fkw["h"])
        E = tb["edge"]
        rret, rkps, rdesc = ref.extract(img, lap, True, taps=True)
        oret, okps, odesc = orc.extract(img, lap, True)
        sizes = [(int(a), int(b)) for a, b in zip(tb["level_w"], tb["level_h"])]
        rows_ok = in_bounds_rows(rkps, sizes, tb["scale"]) if E < 19 else np.ones(len(rkps), bool)
        assert rret == oret and rkps.tobytes() == okps.tobytes(), name
        assert np.array_equal(rdesc[rows_ok], odesc[rows_ok]), name
        assert list(orc.features_per_level()) == list(tb["features_per_level"]) and orc.edge == E
        assert [np.array_equal(a, b) for a, b in zip(orc.scale_factors(), (tb["scale"], tb["inv_scale"], tb["sigma2"], tb["inv_sigma2"]))] == [True] * 4
        assert np.array_equal(orc.umax(), tb["umax"])
        for l in range(okw["nlevels"]):
            assert np.array_equal(orc.level(l), ref.last["levels"][l]), (name, "level", l)
            if ref.last["blurred"][l] is not None:
                assert np.array_equal(orc.blurred(l), ref.last["blurred"][l]), (name, "blur", l)
        ncand = sum(len(orc.candidates(l)[0]) for l in range(okw["nlevels"]))
        assert ncand == ref.last["candidates"], (name, ncand, ref.last["candidates"])
        # keypoints-only overload
        kret, kkps, _ = R.RefOrb(*args).extract(img, lap, False)
        assert kret == rret and kkps.tobytes() == rkps.tobytes()
        # heap-policy and FMA sensitivity
        mret, mkps, mdesc = R.RefOrb(*args, alloc_mode=R.ALLOC_MALLOC).extract(img, lap, True)
        a = kpset(rkps); b = kpset(mkps)
        rep = dict(n=len(rkps), ret=int(rret), edge=int(E), fast_calls=ref.last["fast_calls"], candidates=ref.last["candidates"],
                   desc_rows_in_bounds=int(rows_ok.sum()), glibc_malloc_n=len(mkps), glibc_malloc_kps_not_in_pin=len(b - a),
                   pin_kps_not_in_glibc_malloc=len(a - b))
        if os.path.exists(R.FMA_LIB_PATH) or R.can_build():
            try:
                R.build("ref-fma")
                fret, fkps, fdesc = R.RefOrb(*args, lib_path=R.FMA_LIB_PATH).extract(img, lap, True)
                rep["fma_build_same_keypoints"] = bool(fkps.tobytes() == rkps.tobytes())
                if rep["fma_build_same_keypoints"]:
                    rep["fma_build_desc_bytes_differ"] = int((fdesc[rows_ok] != rdesc[rows_ok]).sum())
                    rep["fma_build_desc_rows_differ"] = int((fdesc[rows_ok] != rdesc[rows_ok]).any(axis=1).sum())
            except Exception as e:   # x86-64-v3 not runnable on this host
                rep["fma_build"] = "unavailable: %r" % (e,)
        report["cases"][name] = rep
        np.savez_compressed(os.path.join(HERE, "ref_orb_%s.npz" % name), frame_kw=str(fkw), orb_kw=str(okw), lapping=np.array(lap, np.int32),
                            ret=np.array([rret], np.int32), kps=rkps, desc=rdesc, desc_rows_defined=rows_ok, edge=np.array([E], np.int32),
                            features_per_level=tb["features_per_level"], scale=tb["scale"], inv_scale=tb["inv_scale"], sigma2=tb["sigma2"],
                            inv_sigma2=tb["inv_sigma2"], umax=tb["umax"], level_w=tb["level_w"], level_h=tb["level_h"],
                            level_sha=np.array([sha(x) for x in ref.last["levels"]]),
                            blur_sha=np.array([sha(x) if x is not None else "" for x in ref.last["blurred"]]),
                            fast_calls=np.array([ref.last["fast_calls"]], np.int64), candidates=np.array([ref.last["candidates"]], np.int64))
        print("golden %-20s n=%4d ret=%4d  oracle == libref;  glibc-malloc heap: %d/%d keypoints differ" %
              (name, len(rkps), rret, len(a - b), len(rkps)))

    # ------------------------------------------------------------------ 2. extractor fuzz (digests)
    rng = np.random.default_rng(20261018)
    cases, rets, ns, ksha, dsha = [], [], [], [], []
    heap_diff = []
    for i in range(N_FUZZ):
        c = fuzz_case(rng)
        img = fuzz_frame(c)
        args = (c["nfeat"], c["sf"], c["nlev"], c["ini"], c["mn"], c["edge"], c["w"], c["h"])
        lap = (c["lap0"], c["lap1"])
        rret, rkps, rdesc = R.RefOrb(*args).extract(img, lap, bool(c["want"]))
        oret, okps, odesc = O.OrbOracle(*args).extract(img, lap, bool(c["want"]))
        assert rret == oret and rkps.tobytes() == okps.tobytes(), ("fuzz", i, c)
        assert (not c["want"]) or np.array_equal(rdesc, odesc), ("fuzz desc", i, c)
        if i < 40:
            _, mkps, _ = R.RefOrb(*args, alloc_mode=R.ALLOC_MALLOC).extract(img, lap, False)
            a = kpset(rkps); b = kpset(mkps)
            heap_diff.append((len(a - b), len(rkps)))
        cases.append(json.dumps(c)); rets.append(rret); ns.append(len(rkps)); ksha.append(sha(rkps))
        dsha.append(sha(rdesc) if c["want"] else "")
    np.savez_compressed(os.path.join(HERE, "ref_orb_fuzz.npz"), cases=np.array(cases), ret=np.array(rets, np.int32), n=np.array(ns, np.int32),
                        kps_sha=np.array(ksha), desc_sha=np.array(dsha))
    report["fuzz"] = dict(cases=N_FUZZ, keypoints_total=int(sum(ns)), oracle_equals_libref=True,
                          glibc_malloc_keypoints_differ=[int(sum(a for a, _ in heap_diff)), int(sum(b for _, b in heap_diff))])
    print("fuzz: %d cases, %d keypoints, oracle == libref on all; glibc-malloc heap: %d/%d keypoints differ over the first %d cases" %
          (N_FUZZ, sum(ns), sum(a for a, _ in heap_diff), sum(b for _, b in heap_diff), len(heap_diff)))

    # ------------------------------------------------------------------ 3. DistributeOctTree alone (ties on purpose)
    rng = np.random.default_rng(7)
    oc, out_sha, out_n = [], [], []
    nties = 0
    for i in range(N_OCTREE):
        c = octree_case(rng)
        r = R.distribute_octtree(c["x"], c["y"], c["resp"], 16, 16 + c["w"], 16, 16 + c["h"], c["N"])
        o = O.distribute_octtree(c["x"], c["y"], c["resp"], 16, 16 + c["w"], 16, 16 + c["h"], c["N"])
        assert np.array_equal(r, o), ("octree", i, c["w"], c["h"], c["N"], len(c["x"]))
        m = R.distribute_octtree(c["x"], c["y"], c["resp"], 16, 16 + c["w"], 16, 16 + c["h"], c["N"], alloc_mode=R.ALLOC_MALLOC)
        nties += int(not np.array_equal(np.sort(m), np.sort(r)))
        oc.append(c); out_sha.append(sha(r)); out_n.append(len(r))
    np.savez_compressed(os.path.join(HERE, "ref_octree_fuzz.npz"), w=np.array([c["w"] for c in oc], np.int32), h=np.array([c["h"] for c in oc], np.int32),
                        N=np.array([c["N"] for c in oc], np.int32), start=np.cumsum([0] + [len(c["x"]) for c in oc]).astype(np.int64),
                        x=np.concatenate([c["x"] for c in oc]), y=np.concatenate([c["y"] for c in oc]), resp=np.concatenate([c["resp"] for c in oc]),
                        out_sha=np.array(out_sha), out_n=np.array(out_n, np.int32))
    report["octree"] = dict(cases=N_OCTREE, oracle_equals_libref=True, cases_where_glibc_malloc_selects_a_different_set=nties)
    print("octree: %d cases, oracle == libref on all; %d cases select a different set under glibc malloc" % (N_OCTREE, nties))

    # ------------------------------------------------------------------ 4. secondary API + DescriptorDistance
    img = synth.make_frame(5)
    ref = R.RefOrb(); orc = O.OrbOracle()
    _, kps, desc = ref.extract(img)
    sel = kps[::7].copy()
    td_r = ref.tracked_desc(img, sel); td_o = orc.tracked_desc(img, sel)
    assert np.array_equal(td_r, td_o)
    moved = np.roll(img, (2, 3), axis=(0, 1))
    al_r = ref.assign_level_by_best_desc(td_r, moved, sel); al_o = orc.assign_level_by_best_desc(td_o, moved, sel)
    assert al_r.tobytes() == al_o.tobytes()
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, (64, 32), dtype=np.uint8); db = rng.integers(0, 256, (512, 32), dtype=np.uint8)
    db[:40] = q[:40] ^ (rng.integers(0, 256, (40, 32), dtype=np.uint8) & rng.integers(0, 256, (40, 32), dtype=np.uint8) & 3)
    D = R.descriptor_distance_matrix(q, db)
    assert all(O.descriptor_distance(q[i], db[j]) == D[i, j] for i in range(0, 64, 9) for j in range(0, 512, 37))
    np.savez_compressed(os.path.join(HERE, "ref_secondary.npz"), frame_seed=np.array([5]), sel=sel, tracked_desc=td_r, assigned=al_r,
                        q=q, db=db, dist=D.astype(np.int16))
    print("secondary API + DescriptorDistance: oracle == libref")

    # ------------------------------------------------------------------ 5. event frames
    ev_cases = []
    K240 = np.array([199.1, 198.9, 132.2, 110.7], np.float32); K346 = np.array([226.4, 226.4, 173.6, 133.7], np.float32)

    def rot(rx, ry, rz):
        cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
        return (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @
                np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))

    def T(r, t):
        m = np.eye(4, dtype=np.float32); m[:3, :3] = rot(*r); m[:3, 3] = t
        return m

    specs = [
        dict(n=2000, seed=3, w=240, h=180, sigma=1.0, mode=0, pol=0), dict(n=2000, seed=3, w=240, h=180, sigma=1.0, mode=0, pol=1),
        dict(n=2000, seed=3, w=240, h=180, sigma=1.0, mode=1, pol=0), dict(n=2000, seed=4, w=240, h=180, sigma=1.0, mode=1, pol=1),
        dict(n=6000, seed=5, w=240, h=180, sigma=0.5, mode=1, pol=0), dict(n=3000, seed=6, w=346, h=260, sigma=2.0, mode=1, pol=0),
        dict(n=6000, seed=7, w=240, h=180, sigma=1.0, mode=2, pol=0, Tcw=T((0.02, -0.03, 0.05), (0.01, -0.02, 0.005)), depth=1.5, K=K240),
        dict(n=50000, seed=8, w=346, h=260, sigma=1.0, mode=2, pol=0, Tcw=T((-0.04, 0.02, 0.08), (0.0, 0.0, 0.0)), depth=1.0, K=K346),
        dict(n=4000, seed=9, w=240, h=180, sigma=1.0, mode=2, pol=1, Tcw=T((3.0, 0.1, 0.2), (0.05, 0.0, -0.02)), depth=2.0, K=K240),   # trace <= 0 branch
        dict(n=6000, seed=10, w=240, h=180, sigma=1.0, mode=3, pol=0, se2=[0.05, 1.0, -2.0], K=K240),
        dict(n=6000, seed=11, w=346, h=260, sigma=1.0, mode=3, pol=0, se2=[-0.03, 2.0, 1.0, 0.9], K=K346),
        dict(n=0, seed=12, w=240, h=180, sigma=1.0, mode=1, pol=0),
    ]
    ev_out = {}
    for i, s in enumerate(specs):
        ev = synth.make_events(s["n"], s["seed"], s["w"], s["h"]) if s["n"] else np.zeros(0, synth.make_events(1, 0).dtype)
        kw = dict(sigma=s["sigma"], mode=s["mode"], pol=bool(s["pol"]), Tcw=s.get("Tcw"), depth=s.get("depth", 1.0), K=s.get("K"), se2=s.get("se2"))
        rf, ru = R.ev_accumulate(ev, s["w"], s["h"], normalize=True, **kw)
        of, mm, ou = O.ev_accumulate(ev, s["w"], s["h"], normalize=True, **kw)
        assert rf.tobytes() == of.tobytes(), ("event float frame", i)
        assert (ru is None) == (ou is None) and (ru is None or np.array_equal(ru, ou)), ("event u8 frame", i)
        ev_out.setdefault("f_sha", []).append(sha(rf)); ev_out.setdefault("u_sha", []).append(sha(ru) if ru is not None else "")
        if i in FULL_EVENT_FRAMES:      # full float frames only for a few cases (fixture size); digests for all
            ev_out["f%d" % i] = rf
        if ru is not None:
            ev_out["u%d" % i] = ru
        # focus measures on the frame
        for what in range(4):
            for avg in (1, 0):
                a = R.lib().ref_image_focus(R._p(rf), s["w"], s["h"], what, avg)
                ev_out.setdefault("focus%d" % i, []).append(a)
        ev_cases.append(json.dumps({k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in s.items()}))
    ev_out["f_sha"] = np.array(ev_out["f_sha"]); ev_out["u_sha"] = np.array(ev_out["u_sha"])
    for i in range(len(specs)):
        ev_out["focus%d" % i] = np.array(ev_out["focus%d" % i], np.float32)
    # motion-compensation Jacobian
    ev = synth.make_events(6000, 21)
    Rt = np.concatenate([rot(0.01, -0.02, 0.03).reshape(9), [0.02, -0.01, 0.005]]).astype(np.float64)
    for glob in (0, 1):
        jr = np.zeros(6); jo = np.zeros(6)
        R.lib().ref_ev_mci_jac(R._p(ev), len(ev), 240, 180, 1.0, R._p(Rt), 1.5, R._p(K240), 0, glob, R._p(jr))
        O.lib().orc_ev_mci_jac(O._p(ev), len(ev), 240, 180, 1.0, O._p(Rt), 1.5, O._p(K240), 0, glob, O._p(jo))
        assert np.abs(jr - jo).max() <= 1e-9 * max(1.0, np.abs(jr).max()), ("jac", glob, jr, jo)
        ev_out["jac%d" % glob] = jr
    np.savez_compressed(os.path.join(HERE, "ref_events.npz"), cases=np.array(ev_cases), jac_Rt=Rt, jac_K=K240, **ev_out)
    report["events"] = dict(cases=len(specs), float_frames_bitwise_equal=True, u8_frames_equal=True, jacobian_max_rel_diff="<= 1e-9")
    print("events: %d cases, oracle float frames == libref bit for bit, normalised u8 equal, Jacobian equal" % len(specs))

    with open(os.path.join(HERE, "ref_report.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print("wrote tests/golden/ref_*.npz and ref_report.json")


if __name__ == "__main__":
    main()
