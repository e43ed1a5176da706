"""The C++ shims with the reference's class signatures (eorb_slam_b200/shim/) compile against a minimal cv mock,
link against libeorb_b200.so and behave like the reference surface.  CPU: loud failure, no fallback.
GPU: results equal the oracle's on the same image."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "eorb_slam_b200", "shim")
EXE = os.path.join(SHIM, "_shim_selftest")


def _build():
    srcs = [os.path.join(SHIM, f) for f in ("shim_selftest.cc", "ORBextractor.cc", "ORBmatcher_b200.cc", "EventConversion_b200.cc", "KLT_b200.cc", "ORBmatcher_guided_b200.cc", "ORBmatcher_kf_b200.cc", "ORBVocabulary_b200.cc")]
    cmd = ["g++", "-std=c++17", "-O1", "-DEORB_SHIM_MOCK", "-I" + os.path.join(SHIM, "cv_mock"), "-I" + SHIM,
           "-I" + os.path.join(ROOT, "include"), "-pthread", "-o", EXE] + srcs + ["-L" + os.path.join(ROOT, "eorb_slam_b200"), "-leorb_b200",
                                                                      "-Wl,-rpath," + os.path.join(ROOT, "eorb_slam_b200")]
    subprocess.check_call(cmd)


def _run():
    from eorb_slam_b200 import api  # noqa: F401  (makes sure the library exists)
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout, r.stderr


def _lcg_image():
    W, H = 752, 480
    s = 12345
    out = np.empty((H, W), np.uint8)
    for y in range(H):
        row = out[y]
        for x in range(W):
            s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
            row[x] = (((x // 24 + y // 24) & 1) * 90 + 60 + (s >> 28)) & 0xFF
    return out


def test_shims_compile_and_fail_loudly_without_device():
    from eorb_slam_b200 import api
    out, err = _run()
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert "ret=-1 n=0" in out and "empty_ret=-1" in out
    assert "no CUDA device" in err
    assert "lk_ok=0 lk_n=0" in out and "elk_n=0 elk_nm1=0" in out
    assert "sfi_nm=0 sfi_self=0" in out and "sbp_nm=0 sbp_set=0" in out and "slp_nm=0 slp_set=0" in out and "sbb_nm=0 sbb_set=0" in out and "sbk_nm=0 sbk_set=0" in out
    assert "sbs_nm=0 sbs_set=0" in out and "sbr_nm=0 sbr_set=0" in out
    assert "kfp_nm=0 kfp_same=0 kfq_nm=0" in out and "fuse_n=0 fuse_add=0 fuse2_n=0" in out and "sim3_n=0" in out and "tri_nm=0 tri_pairs=0" in out
    assert "voc_ok=0" in out and "undist_ok=0" in out


@pytest.mark.gpu
def test_shims_match_oracle_on_gpu():
    import oracle_lib as O
    out, err = _run()
    m = re.search(r"ret=(-?\d+) n=(\d+) desc_rows=(\d+) ret2=(-?\d+) n2=(\d+) levels=(\d+) pyr0=(\d+)x(\d+)", out)
    assert m, out + err
    ret, n, drows, ret2, n2, levels, pw, ph = map(int, m.groups())
    oret, okps, odesc = O.OrbOracle().extract(_lcg_image())
    assert (ret, n, drows, ret2, n2, levels, pw, ph) == (oret, len(okps), len(okps), oret, len(okps), 8, 752, 480)
    assert "empty_ret=-1" in out
    assert "threads_equal=1" in out, "two threads on one extractor object (a device handle per calling thread)"
    d01 = int(re.search(r"dist01=(\d+)", out).group(1))
    assert d01 == O.descriptor_distance(odesc[0], odesc[1])
    sm = re.search(r"selfmatch=(\d+) of (\d+)", out)
    assert int(sm.group(2)) == len(okps) and int(sm.group(1)) > 0.5 * len(okps)
    ev = re.search(r"ev_sum=([\d.]+) ev_u8_max=(\d+) types=(\d+),(\d+) mci_sum=([\d.]+)", out)
    assert int(ev.group(2)) == 255 and (int(ev.group(3)), int(ev.group(4))) == (5, 0)     # CV_32FC1, CV_8UC1
    assert abs(float(ev.group(1)) - float(ev.group(5))) < 1e-2 * float(ev.group(1))        # identity pose == plain splat
    lk = re.search(r"lk_ok=(\d+) lk_n=(\d+) lk_tracked=(\d+) lk_good=(\d+)", out)
    assert lk and int(lk.group(1)) == 1 and int(lk.group(2)) == len(okps)
    assert int(lk.group(3)) > 0.8 * len(okps) and int(lk.group(4)) > 0.9 * int(lk.group(3))      # the (2, 1) shift is recovered
    ek = re.search(r"elk_n=(\d+) elk_nm1=(\d+) elk_good1=(\d+) elk_disp=(\d+) elk_nm2=(\d+) elk_lvl0=(\d+) elk_cnt3=(\d+) elk_last=(\d+)", out)
    en, nm1, good1, ndisp, nm2, lvl0, cnt3, nlast = map(int, ek.groups())
    assert en == len(okps) == nlast and 0 <= int(lk.group(3)) - nm1 <= 5 and 0 <= int(lk.group(4)) - good1 <= 5   # first frame == the plain LK call (minus points outside the image)
    assert 0 < nm2 <= lvl0 < nm1 and ndisp >= lvl0 and 0.9 * lvl0 < cnt3 <= lvl0                    # Init: only octave-0 references stay matched
    fo = re.search(r"focus=([\d.]+) focus_med=([\d.]+) focus_glob=([\d.]+) mean_loc=([\d.]+)", out)
    fv = [float(fo.group(k)) for k in range(1, 5)]
    assert fv[0] > 0 and fv[1] > 0 and fv[2] > fv[0] * 0.5 and abs(fv[3] * 240 * 180 - float(ev.group(1))) < 0.02 * float(ev.group(1))   # mean x pixels = sum
    jm = re.search(r"jac_ok=(\d) jac=([-\d.e+,]+)", out)
    jv = [float(x) for x in jm.group(2).split(",")]
    assert int(jm.group(1)) == 1 and any(abs(v) > 0 for v in jv[3:])
    sf = re.search(r"sfi_nm=(\d+) sfi_self=(\d+) sfi_lvl0=(\d+) sfi_prev=(\d+)", out)
    nm, self_, lvl0, prev = map(int, sf.groups())
    # only level-0 keypoints are queried; a frame matched against its own shifted keypoints finds (nearly) all of them
    assert 0 < nm <= lvl0 and self_ > 0.8 * nm and prev == nm
    vo = re.search(r"voc_ok=(\d) voc_words=(\d+) bow=(\d+) bow_sum=([\d.]+) fv_nodes=(\d+) fv_feats=(\d+)", out)
    okv, words, nbow, bsum, fvn, fvf = vo.groups()
    assert int(okv) == 1 and int(words) == 9 and 0 < int(nbow) <= 9 and abs(float(bsum) - 1.0) < 1e-6      # L1-normalised BowVector
    assert 0 < int(fvn) <= 3 and int(fvf) == len(okps)                                                    # every feature under one of 3 nodes
    un = re.search(r"undist_ok=(\d) undist_n=(\d+) undist_shift=([\d.]+)", out)
    assert int(un.group(1)) == 1 and int(un.group(2)) == len(okps) and float(un.group(3)) > len(okps)      # EuRoC distortion moves points by pixels
    sb = re.search(r"sbp_nm=(\d+) sbp_set=(\d+) sbp_same=(\d+)", out)
    nmp, nset, nsame = map(int, sb.groups())
    # every map point projects exactly onto its keypoint in the current frame: (nearly) all are found, and at their own index
    assert nmp == nset and nmp > 0.9 * len(okps) and nsame > 0.95 * nset
    sl = re.search(r"slp_nm=(\d+) slp_set=(\d+) slp_same=(\d+)", out)
    nlp, lset, lsame = map(int, sl.groups())
    # SearchByProjection(F, vpMapPoints, ...): all points have observations, so nmatches = slots set; most at their own keypoint
    assert nlp == lset and nlp > 0.8 * len(okps) and lsame > 0.9 * lset
    ss = re.search(r"sbs_nm=(\d+) sbs_set=(\d+)", out)
    # rectified stereo with consistent right columns, forward level window [octave, inf): nearly the monocular result
    assert int(ss.group(1)) == int(ss.group(2)) and int(ss.group(1)) > 0.9 * len(okps)
    sr = re.search(r"sbr_nm=(\d+) sbr_set=(\d+) sbr_same=(\d+) sbr_skipped_found=(\d)", out)
    nr, rset, rsame, skipped = map(int, sr.groups())
    # relocalisation search from a keyframe holding the same points: one point is in sAlreadyFound and must not come back
    assert nr == rset and nr > 0.9 * len(okps) and rsame > 0.95 * rset and skipped == 1
    kf = re.search(r"kfp_nm=(\d+) kfp_same=(\d+) kfq_nm=(\d+) kfq_kf=(\d+)", out)
    k1n, k1same, k2n, k2kf = map(int, kf.groups())
    # keyframe-side SearchByProjection (both overloads): every point finds its own keypoint (distance 0), one point per keypoint
    assert k1n == k2n == k2kf and k1n > 0.9 * len(okps) and k1same > 0.95 * k1n
    fu = re.search(r"fuse_n=(\d+) fuse_add=(\d+) fuse2_n=(\d+) fuse2_rep=(\d+)", out)
    f1n, f1add, f2n, f2rep = map(int, fu.groups())
    # Fuse into an empty keyframe adds, into an occupied one proposes replacements; the non-blocking core gives both the same matches
    assert f1n > 0.9 * len(okps) and f2n == f2rep == f1n and 0 < f1add <= f1n
    s3 = re.search(r"sim3_n=(\d+) sim3_same=(\d+)", out)
    s3n, s3same = map(int, s3.groups())
    assert s3n == s3same and s3n > 0.9 * len(okps)      # both directions agree on the point itself
    tr = re.search(r"tri_nm=(\d+) tri_pairs=(\d+) tri_same=(\d+)", out)
    tn, tp, tsame = map(int, tr.groups())
    assert tn == tp and tn > 0.8 * len(okps) and tsame > 0.95 * tn   # SearchForTriangulation: features on their epipolar lines match themselves
    bb = re.search(r"sbb_nm=(\d+) sbb_set=(\d+) sbb_self=(\d+)", out)
    nbb, bset, bself = map(int, bb.groups())
    # SearchByBoW of a frame against a keyframe with the same features: distance 0 to itself, so every match is the feature itself
    assert nbb == bset == bself and nbb > 0.5 * len(okps)
    kk = re.search(r"sbk_nm=(\d+) sbk_set=(\d+) sbk_self=(\d+)", out)
    nkk, kset, kself = map(int, kk.groups())
    assert nkk == kset == kself == nbb         # keyframe-keyframe form on the same data: the same self matches, indexed by the first keyframe
