"""Golden vectors for the SURVEY 8f rank 3 / 4 rows, written by the INDEPENDENT restatements (pure Python for the guided
matchers and the DBoW2 transform, the cv2 wheel for undistortPoints) — not by the C++ oracle, which the tests then check
against these files together with the CUDA path.  Inputs are regenerated from seeds by eorb_slam_b200.synth.
Run:  python tests/golden/make_golden_next_rows.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from eorb_slam_b200 import synth  # noqa: E402
import test_oracle_guided as G  # noqa: E402
import test_oracle_bow as B  # noqa: E402


def main():
    out = {}
    k1, d1, k2, d2, b = synth.make_keypoint_frame_pair(500, 520, 3)
    prev = np.stack([k1["x"], k1["y"]], 1)
    n, m12, p = G.py_search_init(k1, d1, k2, d2, b, prev, 100, 0.9, True)
    out["sfi_args"] = np.array([500, 520, 3, 100]); out["sfi_n"] = np.array([n]); out["sfi_m12"] = m12; out["sfi_prev"] = p
    c = synth.make_projection_case(500, 520, 21)
    n, mc = G.py_search_by_projection(c, 15.0, True)
    out["sbp_args"] = np.array([500, 520, 21, 15]); out["sbp_n"] = np.array([n]); out["sbp_mc"] = mc
    voc = synth.make_vocabulary(10, 3, 1)
    feats = synth.make_vocabulary_features(voc, 300, 11)
    wid, ww, nd, ids, vals, fv = B.py_transform(voc, feats, 2)
    out["bow_args"] = np.array([10, 3, 1, 300, 11, 2]); out["bow_word"] = np.array(wid, np.uint32); out["bow_node"] = np.array(nd, np.uint32)
    out["bow_ids"] = np.array(ids, np.uint32); out["bow_vals"] = np.array(vals, np.float64)
    out["fv_nodes"] = np.array(list(fv), np.uint32)
    out["fv_start"] = np.cumsum([0] + [len(v) for v in fv.values()]).astype(np.int32)
    out["fv_feats"] = np.array([i for v in fv.values() for i in v], np.uint32)
    import cv2
    K = (458.654, 457.296, 367.215, 248.375); D = (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)
    rng = np.random.default_rng(5)
    pts = np.stack([rng.uniform(0, 752, 500), rng.uniform(0, 480, 500)], 1).astype(np.float32)
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    out["und_K"] = np.array(K, np.float32); out["und_D"] = np.array(D, np.float32); out["und_in"] = pts
    out["und_out"] = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, np.array(D, np.float32), None, Km).reshape(-1, 2)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "next_rows.npz"), **out)
    print("wrote next_rows.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
