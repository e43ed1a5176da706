#!/usr/bin/env python3
"""Stage-by-stage GPU-vs-oracle report (does not abort at the first difference).  Debug aid for the GPU box."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from eorb_slam_b200 import api, synth  # noqa: E402


def orb_report(name, img, okw):
    print("== ORB", name, img.shape, okw)
    p = api.ORBxParams(okw["nfeatures"], okw["scale_factor"], okw["nlevels"], okw["ini_th"], okw["min_th"], okw["edge"], (img.shape[1], img.shape[0]))
    ex = api.ORBextractor(p)
    orc = O.OrbOracle(okw["nfeatures"], okw["scale_factor"], okw["nlevels"], okw["ini_th"], okw["min_th"], okw["edge"], img.shape[1], img.shape[0])
    ret, kps, desc = ex(img)
    oret, okps, odesc = orc.extract(img)
    print(" ret", ret, oret, "n", len(kps), len(okps))
    for l in range(okw["nlevels"]):
        a = ex.pyramid_level(l); b = orc.level(l)
        line = " L%d pyr diff px %d/%d" % (l, int((a != b).sum()), a.size)
        gx, gy, gs = ex.debug_candidates(l); ox, oy, os_ = orc.candidates(l)
        sg = set(zip(gx.tolist(), gy.tolist(), gs.tolist())); so = set(zip(ox.tolist(), oy.tolist(), os_.tolist()))
        line += " | cand gpu %d cpu %d common %d order_ok %s" % (len(gx), len(ox), len(sg & so), len(gx) == len(ox) and bool(np.array_equal(gx, ox) and np.array_equal(gy, oy)))
        kx, ky, ks, ka = ex.debug_level_kps(l); px, py, ps, pa = orc.level_kps(l)
        same = len(kx) == len(px) and bool(np.array_equal(kx, px) and np.array_equal(ky, py))
        line += " | sel gpu %d cpu %d same %s" % (len(kx), len(px), same)
        if same and len(kx):
            line += " angle maxdiff %.3g bits_equal %s" % (float(np.abs(ka - pa).max()), bool(np.array_equal(ka.view(np.uint32), pa.view(np.uint32))))
        ob = orc.blurred(l)
        if ob is not None:
            line += " | blur diff %d" % int((ex.debug_blurred(l) != ob).sum())
        print(line)
    if len(kps) == len(okps):
        print(" kps equal", kps.tobytes() == okps.tobytes(), "desc rows differing", int((desc != odesc).any(axis=1).sum()))
        if kps.tobytes() != okps.tobytes():
            for f in kps.dtype.names:
                print("   field", f, "ndiff", int((kps[f] != okps[f]).sum()))


def main():
    print("devices", api.device_count())
    cfg1 = dict(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, edge=19)
    for name, img, okw in [("cfg1", synth.make_frame(0), cfg1),
                           ("ethz", synth.make_frame(4, 240, 180), dict(nfeatures=1000, scale_factor=1.2, nlevels=4, ini_th=10, min_th=0, edge=9)),
                           ("ev1lvl", synth.make_frame(5, 240, 180), dict(nfeatures=400, scale_factor=1.0, nlevels=1, ini_th=0, min_th=0, edge=9))]:
        try:
            orb_report(name, img, okw)
        except Exception:
            traceback.print_exc()
    try:
        db = synth.make_descriptor_db(5000, 1); q, _ = synth.make_queries(db, 300, 2)
        m = api.ORBmatcher(0.7); m.set_db(db); got = m.search(q); exp = O.hamming_best2(q, db, 50, 0.7)
        print("== MATCH", {k: int((got[k] != exp[k]).sum()) for k in got.dtype.names})
        print("   popc rate %.3e /s" % api.probe_popc_rate(0))
    except Exception:
        traceback.print_exc()
    try:
        ev = synth.make_events(2000, 3); cv = api.EvImConverter(0, 1, 100000, 346, 260)
        f = cv.ev2im_gauss(ev, 240, 180, 1.0, False, False); ref, _, _ = O.ev_accumulate(ev, 240, 180, 1.0, mode=1)
        print("== EVENTS gauss rel err %.3g peak %.4f" % (float(np.abs(f - ref).max()) / float(ref.max()), float(ref.max())))
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
