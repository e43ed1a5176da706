"""GPU parity of the bag-of-words / undistortion row through the C ABI against the CPU oracle:
word / node ids, BowVector doubles and FeatureVector lists bit-exact; undistorted keypoints bit-exact."""
import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _api():
    from eorb_slam_b200 import api
    return api


def _same(got, exp):
    for k in ("word_id", "node_id", "bow_ids", "fv_nodes", "fv_start", "fv_feats"):
        assert np.array_equal(got[k], exp[k]), k
    assert got["bow_vals"].tobytes() == exp["bow_vals"].tobytes()          # doubles, bit for bit


@pytest.mark.parametrize("k,L,seed,scoring,weighting,levelsup,n", [
    (10, 3, 1, 0, 0, 2, 300), (10, 4, 2, 0, 0, 4, 1009), (6, 5, 3, 1, 0, 3, 777), (10, 3, 4, 5, 1, 1, 300), (9, 3, 5, 0, 2, 2, 1),
    (4, 6, 6, 2, 3, 4, 2500), (10, 5, 7, 0, 0, 4, 5024), (20, 2, 8, 0, 0, 1, 8192), (10, 4, 9, 0, 0, 0, 64), (10, 4, 10, 0, 0, 9, 64)])
def test_transform_matches_oracle(k, L, seed, scoring, weighting, levelsup, n):
    api = _api()
    voc = synth.make_vocabulary(k, L, seed, scoring, weighting)
    feats = synth.make_vocabulary_features(voc, n, seed + 10)
    v = api.ORBVocabulary(voc)
    got = v.transform(feats, levelsup)
    exp = O.VocabOracle(voc).transform(feats, levelsup)
    _same(got, exp)
    got2 = v.transform(feats[: n // 2], levelsup)                           # same handle, smaller call: no stale state
    _same(got2, O.VocabOracle(voc).transform(feats[: n // 2], levelsup))


def test_transform_edge_cases():
    api = _api()
    voc = synth.make_vocabulary(10, 3, 11)
    v = api.ORBVocabulary(voc)
    got = v.transform(np.zeros((0, 32), np.uint8), 4)
    assert len(got["bow_ids"]) == 0 and len(got["fv_nodes"]) == 0
    voc0 = dict(voc); voc0["weight"] = np.zeros_like(voc["weight"])         # every word stopped -> empty vectors
    feats = synth.make_vocabulary_features(voc, 100, 3)
    got = api.ORBVocabulary(voc0).transform(feats, 4)
    assert len(got["bow_ids"]) == 0 and len(got["fv_nodes"]) == 0 and np.array_equal(got["word_id"], O.VocabOracle(voc0).transform(feats, 4)["word_id"])
    with pytest.raises(api.EorbError):
        v.transform(np.zeros((api.lib.eorb_version() * 0 + 8193, 32), np.uint8), 4)


def test_extract_then_bow_without_leaving_hbm():
    """ORB descriptors straight from the extractor's device output into the vocabulary (Frame::ExtractORB -> ComputeBoW)"""
    import torch
    api = _api()
    img = synth.make_frame(5)
    ex = api.ORBextractor(api.ORBxParams(), 0, 1)
    cap = ex.cap
    st = torch.cuda.current_stream().cuda_stream
    ex.set_stream(st)
    d_img = torch.from_numpy(img).cuda()
    d_kps = torch.zeros(cap * 28, dtype=torch.uint8, device="cuda"); d_desc = torch.zeros(cap * 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(1, dtype=torch.int32, device="cuda")
    ex.extract_batch_raw(d_img.data_ptr(), 1, 752, 480, 752, 752 * 480, (0, 1000), True, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr(),
                         d_mono.data_ptr(), device=True)
    torch.cuda.synchronize()
    n = int(d_n.item())
    voc = synth.make_vocabulary(10, 4, 12)
    v = api.ORBVocabulary(voc)
    v.set_stream(st)
    got = v.transform_device(d_desc.data_ptr(), n, 4)
    _, _, odesc = O.OrbOracle().extract(img)
    _same(got, O.VocabOracle(voc).transform(odesc, 4))
    v.set_stream(None); ex.set_stream(None)


@pytest.mark.parametrize("K,D", [((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)),
                                 ((226.38, 226.15, 173.65, 133.73), (-0.048, 0.011, -0.0005, 0.0003, 0.002)),
                                 ((199.09, 198.83, 132.19, 110.71), (-0.368, 0.150, -0.0003, -0.0002, 0.0)),
                                 ((458.654, 457.296, 367.215, 248.375), (0.0, 0.07, 0.0, 0.0, 0.0))])          # k1 == 0: unchanged (Frame.cc:807)
def test_undistort_keypoints(K, D):
    api = _api()
    k1, _, _, _, _ = synth.make_keypoint_frame_pair(3000, 10, 13, w=int(2 * K[2]), h=int(2 * K[3]))
    got = api.UndistortKeyPoints(k1, K, D)
    if D[0] == 0.0:
        assert got.tobytes() == k1.tobytes()
        return
    exp = k1.copy()
    xy = O.undistort_points(np.stack([k1["x"], k1["y"]], 1), K, D)
    exp["x"] = xy[:, 0]; exp["y"] = xy[:, 1]
    assert got.tobytes() == exp.tobytes()
    cv2 = pytest.importorskip("cv2")                                                                          # and the library itself
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    ref = cv2.undistortPoints(np.stack([k1["x"], k1["y"]], 1).reshape(-1, 1, 2), Km, np.array(D, np.float32), None, Km).reshape(-1, 2)
    assert np.array_equal(np.stack([got["x"], got["y"]], 1), ref)
