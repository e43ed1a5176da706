"""Distinct handles used concurrently from several host threads, as the reference does (left / right extraction threads
Frame.cc:122-125, ORB || AKAZE MixedFrame.cpp:97-102, tracking vs event L1 / L2 threads EvTrackManager.cpp:62-65, four
concurrent ev2* calls EvImBuilder.cpp:1165-1193).  ctypes releases the GIL during the calls."""
import threading

import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth

pytestmark = pytest.mark.gpu


def test_concurrent_handles_from_host_threads():
    from eorb_slam_b200 import api
    img = synth.make_frame(6)
    ev_img = np.ascontiguousarray(img[:180, :240])
    ev = synth.make_events(2000, 9)
    db = synth.make_descriptor_db(30000, 4)
    q, _ = synth.make_queries(db, 300, 5)
    exp_img = O.OrbOracle().extract(img)
    exp_ini = O.OrbOracle(5000, 1.2, 8, 20, 7, 19, 752, 480).extract(img)
    exp_ev = O.OrbOracle(400, 1.0, 1, 0, 0, 9, 240, 180).extract(ev_img, (0, 1000), False)
    exp_f, _, _ = O.ev_accumulate(ev, 240, 180, 1.0, mode=1)
    exp_m = O.hamming_best2(q, db, 50, 0.7)
    errors = []

    def run(fn):
        try:
            for _ in range(15):
                fn()
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    def w_img():
        ex = api.ORBextractor(api.ORBxParams())
        def f():
            r, k, d = ex(img)
            assert r == exp_img[0] and k.tobytes() == exp_img[1].tobytes() and np.array_equal(d, exp_img[2])
        run(f)

    def w_ini():
        ex = api.ORBextractor(api.ORBxParams(5000, 1.2, 8, 20, 7, 19, (752, 480)))
        def f():
            r, k, d = ex(img)
            assert k.tobytes() == exp_ini[1].tobytes() and np.array_equal(d, exp_ini[2])
        run(f)

    def w_ev():
        ex = api.ORBextractor(api.ORBxParams(400, 1.0, 1, 0, 0, 9, (240, 180)))
        cv = api.EvImConverter(0, 1, 4096, 240, 180)
        def f():
            r, k, _ = ex(ev_img, None, (0, 1000), False)
            assert k.tobytes() == exp_ev[1].tobytes()
            fr = cv.ev2im_gauss(ev, 240, 180, 1.0, False, False)
            assert float(np.abs(fr - exp_f).max()) <= 1e-4 * float(exp_f.max())
        run(f)

    def w_match():
        m = api.ORBmatcher(0.7, True)
        m.set_db(db)
        def f():
            got = m.search(q)
            assert all(np.array_equal(got[k], exp_m[k]) for k in got.dtype.names)
        run(f)

    ts = [threading.Thread(target=t) for t in (w_img, w_ini, w_ev, w_match, w_img)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(300)
    assert not errors, errors[:3]
