"""GPU parity: Hamming best-2 + ratio test against the CPU oracle (bit-exact incl. ties: lowest index wins)."""
import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _api():
    from eorb_slam_b200 import api
    return api


def _same(a, b):
    for k in ("best_dist", "best_idx", "second_dist", "accepted"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("ndb,nq", [(1, 3), (2, 5), (255, 17), (256, 1), (257, 1025), (5000, 2000), (70001, 1500)])
def test_best2_matches_oracle(ndb, nq):
    api = _api()
    db = synth.make_descriptor_db(ndb, seed=ndb)
    q, src = synth.make_queries(db, nq, seed=nq)
    dup_src = src[: max(1, nq // 8)]
    synth.plant_duplicate_rows(db, dup_src[dup_src >= 0], seed=3)     # exact duplicates -> distance ties
    q2, _ = synth.make_queries(db, nq, seed=nq + 1, max_flips=0)      # k=0 queries hit the duplicates
    q = np.concatenate([q, q2[: nq // 4]])
    for ratio in (0.7, 0.9):
        m = api.ORBmatcher(ratio, True)
        m.set_db(db)
        got = m.search(q)
        exp = O.hamming_best2(q, db, th=50, ratio=ratio)
        _same(got, exp)
    got = api.hamming_best2(q, db, 50, 0.7)
    _same(got, O.hamming_best2(q, db, 50, 0.7))


def test_best2_empty_and_tiny():
    api = _api()
    m = api.ORBmatcher(0.7)
    m.set_db(np.zeros((0, 32), np.uint8))
    q = synth.make_descriptor_db(9, 1)
    got = m.search(q)
    assert (got["best_idx"] == -1).all() and (got["best_dist"] == 256).all() and (got["accepted"] == 0).all()
    m.set_db(q[:1])
    got = m.search(q)
    exp = O.hamming_best2(q, q[:1], 50, 0.7)
    _same(got, exp)
    assert got["second_dist"][0] == 256 and got["best_dist"][0] == 0 and got["accepted"][0] == 1


def test_sharded_merge_equals_global_search():
    """config 4 shape on one GPU: 4 row shards searched separately (device API), partials concatenated as an
    all-gather would, merged with the (dist, global index) ordering == one global scan."""
    import torch
    api = _api()
    ndb, nq, shards = 40000, 700, 4
    db = synth.make_descriptor_db(ndb, seed=5)
    q, src = synth.make_queries(db, nq, seed=6)
    # duplicates planted across shard boundaries: the lower global index must win
    db[30000:30050] = db[100:150]; q[:50] = db[100:150]
    exp = O.hamming_best2(q, db, 50, 0.7)
    d_q = torch.from_numpy(q).cuda()
    gathered = torch.zeros(shards * nq * 16, dtype=torch.uint8, device="cuda")
    ms = []
    per = ndb // shards
    stream = torch.cuda.current_stream().cuda_stream
    for s in range(shards):
        m = api.ORBmatcher(0.7)
        m.set_stream(stream)
        m.set_db(db[s * per:(s + 1) * per], index_offset=s * per)
        m.search_device(d_q.data_ptr(), nq, gathered.data_ptr() + s * nq * 16)
        ms.append(m)
    d_out = torch.zeros(nq * 16, dtype=torch.uint8, device="cuda")
    ms[0].merge_device(gathered.data_ptr(), shards, nq, d_out.data_ptr())
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(synth.MATCH_DTYPE)
    _same(got, exp)
    assert (got["best_idx"][:50] == np.arange(100, 150)).all()


def test_frame_to_frame_matching_with_rotation_check():
    api = _api()
    p = api.ORBxParams()
    ex = api.ORBextractor(p)
    f1 = synth.make_frame(51)
    f2 = np.roll(f1, (3, 5), axis=(0, 1))
    _, k1, d1 = ex(f1)
    _, k2, d2 = ex(f2)
    m = api.ORBmatcher(0.9, True)
    n, match12 = m.SearchBruteForce(d1, d2, k1["angle"], k2["angle"])
    exp = O.hamming_best2(d1, d2, 50, 0.9)
    m12 = np.where(exp["accepted"] == 1, exp["best_idx"], -1).astype(np.int32)
    n_ref, m_ref = O.rotation_filter(k1["angle"], k2["angle"], m12)
    assert n == n_ref and np.array_equal(match12, m_ref)
    assert n > 200


def test_popc_probe_reports_a_rate():
    api = _api()
    r = api.probe_popc_rate(0)
    assert 1e11 < r < 1e14


def test_sharded_search_entry_point_single_rank_communicator():
    """eorb_matcher_search_sharded with a one-rank ncclComm_t: scan + ncclAllGather + merge through the C ABI give
    exactly the one-call result (multi-rank behaviour of the merge is covered on the CPU by tests/test_dist_gloo.py
    and on two B200s by bench.py --gpus 2)."""
    import torch
    api = _api()
    db = synth.make_descriptor_db(30000, 5)
    q, _ = synth.make_queries(db, 500, 6)
    m = api.ORBmatcher(0.7, True)
    m.set_db(db, index_offset=1000)
    exp = O.hamming_best2(q, db, 50, 0.7)
    comm = api.nccl_comm_init_rank(1, api.nccl_unique_id(), 0, 0)
    d_q = torch.from_numpy(q).cuda()
    d_out = torch.zeros(len(q) * 16, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    m.search_sharded(d_q.data_ptr(), len(q), comm, 1, d_out.data_ptr())
    m.synchronize()
    got = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=api.MATCH_DTYPE)
    assert np.array_equal(got["best_dist"], exp["best_dist"]) and np.array_equal(got["second_dist"], exp["second_dist"])
    assert np.array_equal(got["accepted"], exp["accepted"])
    assert np.array_equal(got["best_idx"], np.where(exp["best_idx"] >= 0, exp["best_idx"] + 1000, -1))
    api.nccl_comm_destroy(comm)


# ------------------------------------------------------------------------------------------------ tensor-core engine (hamming_tc.cu)
@pytest.mark.parametrize("ndb,nq", [(1, 3), (255, 17), (256, 128), (257, 129), (5000, 2000), (70001, 1500), (300000, 333)])
def test_tensor_engine_matches_oracle(ndb, nq):
    """tcgen05 int8 contraction (dist = popc(q) - dot) == oracle bit for bit: ties (lowest index wins), exact duplicates, partial
    last tiles, query counts that do not fill the 128-row accumulator, databases smaller than one tile"""
    api = _api()
    db = synth.make_descriptor_db(ndb, seed=ndb + 7)
    q, src = synth.make_queries(db, nq, seed=nq + 3)
    dup_src = src[: max(1, nq // 8)]
    synth.plant_duplicate_rows(db, dup_src[dup_src >= 0], seed=4)
    q2, _ = synth.make_queries(db, nq, seed=nq + 1, max_flips=0)
    q = np.concatenate([q, q2[: nq // 4]])
    q[0] = 0; q[-1] = 255                      # popc(q) = 0 and 256: the ends of the dot range
    if ndb > 4:
        db[1] = 255; db[2] = 0
    m = api.ORBmatcher(0.7, True)
    m.set_db(db)
    m.set_engine(m.HAMMING_TENSOR)
    got = m.search(q)
    assert m.last_engine() == m.HAMMING_TENSOR
    _same(got, O.hamming_best2(q, db, th=50, ratio=0.7))
    m.set_engine(m.HAMMING_POPC)
    _same(m.search(q), got)
    assert m.last_engine() == m.HAMMING_POPC


def test_tensor_engine_sharded_merge_and_auto_rule():
    """row shards with index offsets searched on the tensor engine, merged like the all-gather's output; AUTO picks the tensor engine
    only for large searches"""
    import torch
    api = _api()
    ndb, nq, shards = 400000, 700, 4
    db = synth.make_descriptor_db(ndb, seed=15)
    q, _ = synth.make_queries(db, nq, seed=16)
    db[300000:300050] = db[100:150]; q[:50] = db[100:150]     # duplicates across shards: the lower global index wins
    exp = O.hamming_best2(q, db, 50, 0.7)
    d_q = torch.from_numpy(q).cuda()
    gathered = torch.zeros(shards * nq * 16, dtype=torch.uint8, device="cuda")
    per = ndb // shards
    stream = torch.cuda.current_stream().cuda_stream
    ms = []
    for s in range(shards):
        m = api.ORBmatcher(0.7)
        m.set_stream(stream)
        m.set_db(db[s * per:(s + 1) * per], index_offset=s * per)
        m.search_device(d_q.data_ptr(), nq, gathered.data_ptr() + s * nq * 16)
        assert m.last_engine() == m.HAMMING_TENSOR      # AUTO: 700 queries x 100000 rows
        ms.append(m)
    d_out = torch.zeros(nq * 16, dtype=torch.uint8, device="cuda")
    ms[0].merge_device(gathered.data_ptr(), shards, nq, d_out.data_ptr())
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(synth.MATCH_DTYPE)
    _same(got, exp)
    assert (got["best_idx"][:50] == np.arange(100, 150)).all()
    small = api.ORBmatcher(0.7)
    small.set_db(db[:5000])
    small.search(q[:20])
    assert small.last_engine() == small.HAMMING_POPC    # AUTO: small searches stay on the POPC kernel
