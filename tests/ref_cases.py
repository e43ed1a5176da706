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
