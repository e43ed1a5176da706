"""tests/golden/next_rows.npz (written by tests/golden/make_golden_next_rows.py from the independent Python / cv2 restatements):
the C++ oracle (CPU test) and the CUDA path through the C ABI (gpu test) must both reproduce it bit for bit."""
import os

import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "next_rows.npz"))


def _inputs(gold):
    n1, n2, seed, win = (int(v) for v in gold["sfi_args"])
    k1, d1, k2, d2, b = synth.make_keypoint_frame_pair(n1, n2, seed)
    pn1, pn2, pseed, th = (int(v) for v in gold["sbp_args"])
    c = synth.make_projection_case(pn1, pn2, pseed)
    k, L, vseed, nf, fseed, lup = (int(v) for v in gold["bow_args"])
    voc = synth.make_vocabulary(k, L, vseed)
    feats = synth.make_vocabulary_features(voc, nf, fseed)
    return (k1, d1, k2, d2, b, win), (c, float(th)), (voc, feats, lup)


def _check_bow(got, gold):
    assert np.array_equal(got["word_id"], gold["bow_word"]) and np.array_equal(got["node_id"], gold["bow_node"])
    assert np.array_equal(got["bow_ids"], gold["bow_ids"]) and got["bow_vals"].tobytes() == gold["bow_vals"].tobytes()
    assert np.array_equal(got["fv_nodes"], gold["fv_nodes"]) and np.array_equal(got["fv_start"], gold["fv_start"])
    assert np.array_equal(got["fv_feats"], gold["fv_feats"])


def test_oracle_reproduces_the_golden_vectors(gold):
    (k1, d1, k2, d2, b, win), (c, th), (voc, feats, lup) = _inputs(gold)
    n, m12, p = O.search_for_initialization(k1, d1, k2, d2, b, np.stack([k1["x"], k1["y"]], 1), win, 0.9, True)
    assert n == int(gold["sfi_n"][0]) and np.array_equal(m12, gold["sfi_m12"]) and p.tobytes() == gold["sfi_prev"].tobytes()
    n, mc = O.search_by_projection(c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], c["K"],
                                   c["scale_factors"], th, True)
    assert n == int(gold["sbp_n"][0]) and np.array_equal(mc, gold["sbp_mc"])
    _check_bow(O.VocabOracle(voc).transform(feats, lup), gold)
    assert O.undistort_points(gold["und_in"], gold["und_K"], gold["und_D"]).tobytes() == gold["und_out"].tobytes()


@pytest.mark.gpu
def test_cuda_path_reproduces_the_golden_vectors(gold):
    from eorb_slam_b200 import api
    (k1, d1, k2, d2, b, win), (c, th), (voc, feats, lup) = _inputs(gold)
    gm = api.GuidedMatcher(0, 0.9, True)
    n, m12, p = gm.SearchForInitialization(k1, d1, k2, d2, b, np.stack([k1["x"], k1["y"]], 1), win)
    assert n == int(gold["sfi_n"][0]) and np.array_equal(m12, gold["sfi_m12"]) and p.tobytes() == gold["sfi_prev"].tobytes()
    n, mc = gm.SearchByProjection(c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], c["K"],
                                  c["scale_factors"], th)
    assert n == int(gold["sbp_n"][0]) and np.array_equal(mc, gold["sbp_mc"])
    _check_bow(api.ORBVocabulary(voc).transform(feats, lup), gold)
    kp = np.zeros(len(gold["und_in"]), synth.KEYPOINT_DTYPE)
    kp["x"] = gold["und_in"][:, 0]; kp["y"] = gold["und_in"][:, 1]
    un = api.UndistortKeyPoints(kp, gold["und_K"], gold["und_D"])
    assert np.stack([un["x"], un["y"]], 1).tobytes() == gold["und_out"].tobytes()
