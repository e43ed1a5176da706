"""Writes tests/golden/ref_guided_kf.npz: the results of the REFERENCE'S OWN keyframe-side search bodies (src/ORBmatcher.cc :480-593,
:595-712 SearchByProjection(KeyFrame*, Scw, ...), :1407-1617 and :1619-1741 Fuse, :1743-1967 SearchBySim3, + KeyFrame::GetFeaturesInArea /
IsInImage, src/KeyFrame.cc:873-922; cut out at build time into oracle/_ref/libref.so by oracle/Makefile) on the seeded cases of
tests/kf_cases.py.  Needs the reference tree (authoring container only):

    python tests/golden/make_ref_guided_kf_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import kf_cases as KC       # noqa: E402
import oracle_lib as O      # noqa: E402
import ref_lib as R         # noqa: E402


def main():
    assert R.can_build(), "needs the reference tree (set EORB_REFERENCE)"
    R.build("ref")
    out, total = {}, 0
    for key, kind, ov, c in KC.kf_cases():
        r = KC.run_ref(R, kind, ov, c)
        o = KC.run_composed(O.search_windows, kind, ov, c)
        assert r[0] == o[0] and r[1].shape == o[1].shape and np.array_equal(r[1], o[1]), ("host part + oracle core != libref", key)
        out[key + "_n"] = np.array([r[0]], np.int32); out[key] = r[1].astype(np.int32)
        total += int(r[0])
    # SearchForTriangulation (:975-1214): match table + what the function forms inside and the device path gets from its caller (F12, epipole)
    tri, ntri = {}, 0
    for key, a, kw in KC.tri_cases():
        r = R.search_for_triangulation(*a, **kw)
        o = KC.tri_compose(O.search_for_triangulation, a, kw, r[2], r[3])
        assert r[0] == o[0] and np.array_equal(r[1], o[1]), ("oracle != libref", key)
        tri[key + "_n"] = np.array([r[0]], np.int32); tri[key] = r[1].astype(np.int32); tri[key + "_F"] = r[2]; tri[key + "_ep"] = r[3]
        ntri += int(r[0])
    np.savez_compressed(os.path.join(HERE, "ref_triangulation.npz"), **tri)
    print("ref_triangulation.npz: %d cases, %d matches, oracle == libref on every one" % (sum(k.endswith("_n") for k in tri), ntri))
    np.savez_compressed(os.path.join(HERE, "ref_guided_kf.npz"), **out)
    print("ref_guided_kf.npz: %d cases, %d matches / fusions, host part + oracle core == libref on every one" % (sum(k.endswith("_n") for k in out), total))


if __name__ == "__main__":
    main()
