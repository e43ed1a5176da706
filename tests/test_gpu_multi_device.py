"""One process, two devices: every handle type created on device 1 after device 0 (kernel attributes such as the dynamic
shared-memory limits are per device).  Skipped on single-GPU boxes."""
import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth

pytestmark = pytest.mark.gpu


def test_handles_on_a_second_device():
    from eorb_slam_b200 import api
    if api.device_count() < 2:
        pytest.skip("needs two GPUs")
    img = synth.make_frame(4)
    p = api.ORBxParams(5000, 1.2, 8, 20, 7, 19, (752, 480))
    _, ok, od = O.OrbOracle(5000, 1.2, 8, 20, 7, 19, 752, 480).extract(img)
    k1, d1, k2, d2, b = synth.make_keypoint_frame_pair(5000, 5000, 7)
    prev = np.stack([k1["x"], k1["y"]], 1)
    en, em12, ep = O.search_for_initialization(k1, d1, k2, d2, b, prev, 100, 0.9, True)
    voc = synth.make_vocabulary(10, 3, 1)
    feats = synth.make_vocabulary_features(voc, 5000, 2)
    ebw = O.VocabOracle(voc).transform(feats, 2)
    db = synth.make_descriptor_db(20000, 1)
    q, _ = synth.make_queries(db, 256, 2)
    eh = O.hamming_best2(q, db, 50, 0.7)
    ev = synth.make_events(2000, 3)
    ef, _, _ = O.ev_accumulate(ev, 240, 180, 1.0, mode=1)
    for dev in (0, 1, 0):
        _, k, d = api.ORBextractor(p, dev, 1)(img)
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od), dev
        n, m12, pm = api.GuidedMatcher(dev, 0.9, True).SearchForInitialization(k1, d1, k2, d2, b, prev, 100)
        assert n == en and np.array_equal(m12, em12), dev
        bw = api.ORBVocabulary(voc, dev).transform(feats, 2)
        assert np.array_equal(bw["bow_ids"], ebw["bow_ids"]) and bw["bow_vals"].tobytes() == ebw["bow_vals"].tobytes(), dev
        m = api.ORBmatcher(0.7, True, dev)
        m.set_db(db)
        got = m.search(q)
        assert all(np.array_equal(got[kk], eh[kk]) for kk in got.dtype.names), dev
        f = api.EvImConverter(dev, 1, 4096, 240, 180).ev2im_gauss(ev, 240, 180, 1.0, False, False)
        assert float(np.abs(f - ef).max()) <= 1e-4 * float(ef.max()), dev
