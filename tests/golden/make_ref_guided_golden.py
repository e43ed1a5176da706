"""Writes tests/golden/ref_guided.npz: the results of the REFERENCE'S OWN tracking-thread matcher bodies (src/ORBmatcher.cc :44-218,
:276-478, :714-831, :1969-2187, :2189-2312 + Frame::GetFeaturesInArea, cut out at build time into oracle/_ref/libref.so by
oracle/Makefile) on the seeded cases of tests/ref_cases.py::guided_cases.  Needs the reference tree (authoring container only):

    python tests/golden/make_ref_guided_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O      # noqa: E402
import ref_cases as RC      # noqa: E402
import ref_lib as R         # noqa: E402


def main():
    assert R.can_build(), "needs the reference tree (set EORB_REFERENCE)"
    R.build("ref")
    out, total = {}, 0
    for key, kind, a, kw in RC.guided_cases():
        r = RC.guided_run(R, kind, a, kw)
        o = RC.guided_run(O, kind, a, kw)
        assert r[0] == o[0] and np.array_equal(r[1], o[1]), ("oracle != libref", key)
        if kind == "init":
            assert r[2].tobytes() == o[2].tobytes(), key
            out[key + "_prev"] = r[2]
        out[key + "_n"] = np.array([r[0]], np.int32); out[key] = r[1].astype(np.int32)
        total += int((r[1] >= 0).sum())
    np.savez_compressed(os.path.join(HERE, "ref_guided.npz"), **out)
    print("ref_guided.npz: %d cases, %d matches, oracle == libref on every one" % (sum(k.endswith("_n") for k in out), total))


if __name__ == "__main__":
    main()
