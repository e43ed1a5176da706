"""Shared readers of the reference-pinned fixtures tests/golden/ref_*.npz (written by tests/golden/make_ref_golden.py from
oracle/_ref/libref.so = the reference's own ORBextractor.cc / EventConversion.cc compiled unmodified)."""
import ast
import hashlib
import json
import os

import numpy as np

from eorb_slam_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ORB_NAMES = ["cfg1_seed0", "cfg1_seed1_stereo", "cfg1_flat", "mvsec_346x260", "ethz_240x180_e9", "ev_single_level"]


def sha(a) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def orb_golden(name):
    g = np.load(os.path.join(GOLDEN, "ref_orb_%s.npz" % name))
    fkw = ast.literal_eval(str(g["frame_kw"])); okw = ast.literal_eval(str(g["orb_kw"]))
    return g, fkw, okw, synth.make_frame(**fkw), tuple(int(v) for v in g["lapping"])


def orb_args(okw, fkw):
    return (okw["nfeatures"], okw["scale_factor"], okw["nlevels"], okw["ini_th"], okw["min_th"], okw["edge"], fkw["w"], fkw["h"])


def fuzz_cases():
    g = np.load(os.path.join(GOLDEN, "ref_orb_fuzz.npz"))
    for i, c in enumerate(g["cases"]):
        yield i, json.loads(str(c)), int(g["ret"][i]), int(g["n"][i]), str(g["kps_sha"][i]), str(g["desc_sha"][i])


def fuzz_frame(c):
    return synth.make_frame(c["seed"], c["w"], c["h"], nrect=c["nrect"], noise=c["noise"], kind=c["kind"])


def fuzz_args(c):
    return (c["nfeat"], c["sf"], c["nlev"], c["ini"], c["mn"], c["edge"], c["w"], c["h"])


def octree_cases():
    g = np.load(os.path.join(GOLDEN, "ref_octree_fuzz.npz"))
    st = g["start"]
    for i in range(len(g["w"])):
        a, b = int(st[i]), int(st[i + 1])
        yield i, int(g["w"][i]), int(g["h"][i]), int(g["N"][i]), g["x"][a:b], g["y"][a:b], g["resp"][a:b], str(g["out_sha"][i]), int(g["out_n"][i])


def event_cases():
    g = np.load(os.path.join(GOLDEN, "ref_events.npz"))
    for i, c in enumerate(g["cases"]):
        s = json.loads(str(c))
        ev = synth.make_events(s["n"], s["seed"], s["w"], s["h"]) if s["n"] else np.zeros(0, synth.make_events(1, 0).dtype)
        kw = dict(sigma=s["sigma"], mode=s["mode"], pol=bool(s["pol"]), Tcw=np.array(s["Tcw"], np.float32) if "Tcw" in s else None,
                  depth=s.get("depth", 1.0), K=np.array(s["K"], np.float32) if "K" in s else None, se2=s.get("se2"))
        yield i, s, ev, kw, g


# ----------------------------------------------------------------------------- tracking-thread matchers (ORBmatcher.cc bodies in libref)
def _stereo_u_right(rng, kps, frac=0.6):
    """mvuRight of a rectified pair: a right-image column for `frac` of the features, -1 elsewhere (Frame::ComputeStereoMatches)"""
    n = len(kps)
    return np.where(rng.random(n) < frac, kps["x"] - rng.random(n).astype(np.float32) * 30, -1).astype(np.float32)


def bow_case(n1, n2, seed, k=6, L=3, levelsup=2, valid_frac=0.8, max_flips=24):
    import oracle_lib as O
    k1, d1, k2, d2, _ = synth.make_keypoint_frame_pair(n1, n2, seed, max_flips=max_flips)
    vo = O.VocabOracle(synth.make_vocabulary(k, L, seed))
    t1, t2 = vo.transform(d1, levelsup), vo.transform(d2, levelsup)
    valid = (np.random.default_rng(seed).random(n1) < valid_frac).astype(np.uint8)
    return k1, d1, valid, (t1["fv_nodes"], t1["fv_start"], t1["fv_feats"]), k2, d2, (t2["fv_nodes"], t2["fv_start"], t2["fv_feats"])


def guided_cases(nseeds=8):
    """(key, kind, args, kwargs) of every matcher case pinned in tests/golden/ref_guided.npz.  kinds: 'proj' = SearchByProjection(Cur, Last)
    (ORBmatcher.cc:1969) with level_mode / stereo variants, 'reloc' = SearchByProjection(Cur, pKF, sAlreadyFound) (:2189), 'map' =
    SearchByProjection(F, vpMapPoints) (:44) mono and rectified-stereo, 'init' = SearchForInitialization (:714), 'bow' = SearchByBoW (:276).
    Levels are kept inside [0, nlevels): the reference indexes mvScaleFactors without a clamp (UB outside)."""
    for seed in range(nseeds):
        rng = np.random.default_rng(9000 + seed)
        c = synth.make_projection_case(500, 520, 100 + seed, zero_obs_frac=0.1 * (seed % 4))
        nl = len(c["scale_factors"])
        for mode in (0, 1, 2):
            for stereo in (0, 1):
                ur = _stereo_u_right(rng, c["kps2"]) if stereo else None
                a = (c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], c["K"], c["scale_factors"])
                kw = dict(th=15.0 if seed % 2 else 7.0, check_ori=bool(seed % 3), level_mode=mode, mbf=40.0 if stereo else 0.0, u_right2=ur)
                yield "proj_s%d_m%d_st%d" % (seed, mode, stereo), "proj", a, kw
        lv = np.clip(c["kps1"]["octave"] + rng.integers(-1, 2, len(c["kps1"])), 0, nl - 1).astype(np.int32)
        held = (rng.random(520) < 0.2).astype(np.uint8)
        for od in (100, 64):
            a = (c["x3Dc"], c["valid1"], lv, c["kps1"], c["descMP"], c["kps2"], c["desc2"], held if od == 100 else None, c["bounds"], c["K"],
                 c["scale_factors"])
            yield "reloc_s%d_d%d" % (seed, od), "reloc", a, dict(th=10.0 if od == 100 else 20.0, orb_dist=od, check_ori=bool(seed % 2))
        m = synth.make_local_map_case(600, 520, 200 + seed, zero_obs_frac=0.1 * (seed % 4))
        m["pts"]["scale_level"] = np.clip(m["pts"]["scale_level"], 0, len(m["scale_factors"]) - 1)
        for stereo in (0, 1):
            ur = _stereo_u_right(rng, m["kps2"]) if stereo else None
            xr = (m["pts"]["proj_x"] - rng.random(len(m["pts"])).astype(np.float32) * 30).astype(np.float32) if stereo else None
            a = (m["pts"], xr, m["descMP"], m["kps2"], m["desc2"], m["held2"], ur, m["bounds"], m["scale_factors"])
            kw = dict(th=[1.0, 3.0, 5.0][seed % 3], far_points=bool(seed % 2), th_far=20.0, nnratio=0.8)
            yield "map_s%d_st%d" % (seed, stereo), "map", a, kw
        k1, d1, k2, d2, _ = synth.make_keypoint_frame_pair(500, 520, 300 + seed)
        prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
        yield ("init_s%d" % seed, "init", (k1, d1, k2, d2, np.array([0, 0, 752, 480], np.float32), prev),
               dict(window_size=100 if seed % 2 else 30, nnratio=0.9, check_ori=bool(seed % 3)))
        yield "bow_s%d" % seed, "bow", bow_case(500, 520, 400 + seed, levelsup=seed % 4), dict(nnratio=0.7 if seed % 2 else 0.9, check_ori=bool(seed % 3))


def guided_run(mod, kind, a, kw):
    """run one case through tests/oracle_lib.py (mod = O) or tests/ref_lib.py (mod = R) -> (nmatches, int32 result array[, prev_xy])"""
    import oracle_lib as O
    if kind == "proj":
        return (O.search_by_projection_ex if mod is O else mod.search_by_projection)(*a, **kw)
    if kind == "reloc":
        return mod.search_by_projection_reloc(*a, **kw)
    if kind == "map":
        return (O.search_by_projection_map_points_ex if mod is O else mod.search_by_projection_map_points)(*a, **kw)
    if kind == "init":
        return mod.search_for_initialization(*a, kw["window_size"], kw["nnratio"], kw["check_ori"])
    if kind == "bow":
        return mod.search_by_bow(*a, kw["nnratio"], kw["check_ori"])
    raise KeyError(kind)


def guided_golden():
    return np.load(os.path.join(GOLDEN, "ref_guided.npz"))
