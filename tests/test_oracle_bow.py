"""Pins oracle/bow_oracle.cc: the DBoW2 transform against an independent pure-Python restatement (dict / list based,
written from Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1258, BowVector.cpp, FeatureVector.cpp) and
undistortPoints against the cv2 wheel (the library the reference calls, Frame.cc:805-840)."""
import math

import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth


def py_transform(voc, feats, levelsup):
    n = len(voc["parent"])
    children = [[] for _ in range(n)]
    word = {}
    for nid in range(1, n):
        children[int(voc["parent"][nid])].append(nid)
        if voc["is_leaf"][nid]:
            word[nid] = len(word)
    norm = {0: 1, 1: 2, 2: 1, 3: 1, 4: 1, 5: 0}[voc["scoring"]]
    tf = voc["weighting"] in (0, 1)
    bow, fv = {}, {}
    wid, ww, nd = [], [], []
    for i, f in enumerate(feats):
        node, level, nid = 0, 0, 0
        while True:
            level += 1
            best, bid = None, None
            for c in children[node]:
                d = int(np.unpackbits(np.bitwise_xor(f, voc["desc"][c])).sum())
                if best is None or d < best:
                    best, bid = d, c
            node = bid
            if level == voc["L"] - levelsup:
                nid = node
            if voc["is_leaf"][node]:
                break
        w = float(voc["weight"][node])
        wid.append(word[node]); ww.append(w); nd.append(nid)
        if w > 0:
            if tf:
                bow[word[node]] = bow.get(word[node], 0.0) + w
            else:
                bow.setdefault(word[node], w)
            fv.setdefault(nid, []).append(i)
    ids = sorted(bow)
    vals = [bow[k] for k in ids]
    if tf and ids and norm == 0:
        vals = [v / float(len(ids)) for v in vals]
    if norm:
        s = 0.0
        for v in vals:
            s = s + (abs(v) if norm == 1 else v * v)
        if norm == 2:
            s = math.sqrt(s)
        if s > 0:
            vals = [v / s for v in vals]
    return wid, ww, nd, ids, vals, {k: fv[k] for k in sorted(fv)}


@pytest.mark.parametrize("k,L,seed,scoring,weighting,levelsup", [(10, 3, 1, 0, 0, 2), (10, 4, 2, 0, 0, 4), (6, 5, 3, 1, 0, 3), (10, 3, 4, 5, 1, 1),
                                                                 (9, 3, 5, 0, 2, 2), (4, 6, 6, 2, 3, 4)])
def test_transform_matches_python(k, L, seed, scoring, weighting, levelsup):
    voc = synth.make_vocabulary(k, L, seed, scoring, weighting)
    feats = synth.make_vocabulary_features(voc, 300, seed + 10)
    got = O.VocabOracle(voc).transform(feats, levelsup)
    wid, ww, nd, ids, vals, fv = py_transform(voc, feats, levelsup)
    assert got["word_id"].tolist() == wid and got["word_w"].tolist() == ww and got["node_id"].tolist() == nd
    assert got["bow_ids"].tolist() == ids and got["bow_vals"].tolist() == vals           # doubles, bit for bit
    assert got["fv_nodes"].tolist() == list(fv)
    for q, node in enumerate(fv):
        assert got["fv_feats"][got["fv_start"][q]:got["fv_start"][q + 1]].tolist() == fv[node]
    if scoring == 0 and len(vals):
        assert abs(sum(vals) - 1.0) < 1e-12
    assert len(ids) < len(feats)                                                          # repeated words were merged


def test_transform_empty_inputs():
    voc = synth.make_vocabulary(10, 3, 7)
    got = O.VocabOracle(voc).transform(np.zeros((0, 32), np.uint8), 4)
    assert len(got["bow_ids"]) == 0 and len(got["fv_nodes"]) == 0


@pytest.mark.parametrize("K,D", [((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)),   # EuRoC cam0
                                 ((226.38, 226.15, 173.65, 133.73), (-0.048, 0.011, -0.0005, 0.0003, 0.002)),
                                 ((199.09, 198.83, 132.19, 110.71), (-0.368, 0.150, -0.0003, -0.0002, 0.0))])
def test_undistort_points_matches_cv2(K, D):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    w, h = int(2 * K[2]), int(2 * K[3])
    pts = np.stack([rng.uniform(0, w, 4000), rng.uniform(0, h, 4000)], 1).astype(np.float32)
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    Dm = np.array(D, np.float32)
    exp = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, Dm, None, Km).reshape(-1, 2)
    got = O.undistort_points(pts, K, D)
    assert got.tobytes() == exp.tobytes()
