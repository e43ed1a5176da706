"""ctypes binding of oracle/_ref/libref.so: the REFERENCE'S OWN sources (ORBextractor.cc, EventConversion.cc,
ORBmatcher::DescriptorDistance) compiled unmodified against the stand-in headers of oracle/ref_mock/.

TEST INFRASTRUCTURE ONLY.  libref.so can only be (re)built where /root/reference exists (this container); on the GPU
box the prebuilt file travels with the snapshot, and everything that needs it skips when it is absent.  The committed
fixtures tests/golden/ref_*.npz (tests/golden/make_ref_golden.py) carry its outputs everywhere else.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SRC = os.environ.get("EORB_REFERENCE", "/root/reference")
LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libref.so")
FMA_LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libref_fma.so")

ALLOC_MALLOC, ALLOC_BUMP = 0, 1


def can_build() -> bool:
    return os.path.exists(os.path.join(REF_SRC, "src", "ORBextractor.cc"))


def build(target: str = "ref") -> None:
    """(re)build through the committed recipe; no-op when up to date; needs the reference tree"""
    O.build()
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", target, "REF=" + REF_SRC])


def available() -> bool:
    if can_build():
        try:
            build()
        except Exception:
            return os.path.exists(LIB_PATH)
    return os.path.exists(LIB_PATH)


_libs = {}


def lib(path: str = LIB_PATH):
    if path in _libs:
        return _libs[path]
    if path == LIB_PATH and can_build():
        build()
    O.lib()   # liboracle.so first: libref's primitives forward to it
    L = C.CDLL(path)
    vp = C.c_void_p
    L.ref_version.restype = C.c_int
    L.ref_orb_tables.argtypes = [C.POINTER(O.OrbParams), C.c_int, C.c_int] + [vp] * 9
    L.ref_orb_tables.restype = C.c_int
    L.ref_orb_extract.argtypes = [C.POINTER(O.OrbParams), vp, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                  C.c_int, vp, vp, vp, vp, vp]
    L.ref_orb_extract.restype = C.c_int
    L.ref_distribute_octtree.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    L.ref_distribute_octtree.restype = C.c_int
    L.ref_orb_tracked_desc.argtypes = [C.POINTER(O.OrbParams), vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, vp]
    L.ref_orb_tracked_desc.restype = C.c_int
    L.ref_orb_assign_level_by_best_desc.argtypes = [C.POINTER(O.OrbParams), vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int]
    L.ref_orb_assign_level_by_best_desc.restype = C.c_int
    L.ref_descriptor_distance.argtypes = [vp, vp]; L.ref_descriptor_distance.restype = C.c_int
    L.ref_descriptor_distance_matrix.argtypes = [vp, C.c_int, vp, C.c_int, vp]; L.ref_descriptor_distance_matrix.restype = None
    L.ref_orb_extract_batch_mt.argtypes = [C.POINTER(O.OrbParams), vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.ref_orb_extract_batch_mt.restype = C.c_long
    if hasattr(L, "ref_ev_accumulate"):
        L.ref_ev_accumulate.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, vp, C.c_float, vp, vp, C.c_int, C.c_int, C.c_int,
                                        vp, vp]
        L.ref_ev_accumulate.restype = C.c_int
        L.ref_image_focus.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]; L.ref_image_focus.restype = C.c_float
        L.ref_ev_mci_jac.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_float, vp, C.c_float, vp, C.c_int, C.c_int, vp]
        L.ref_ev_mci_jac.restype = None
        L.ref_ev_accumulate_batch_mt.argtypes = [vp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, vp, C.c_float, vp, C.c_int]
        L.ref_ev_accumulate_batch_mt.restype = C.c_double
    _libs[path] = L
    return L


_p = O._p


class RefOrb:
    """The reference's ORBextractor (ORBextractor.cc, unmodified) behind the same Python surface as oracle_lib.OrbOracle"""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, edge_th=19, im_w=752, im_h=480,
                 alloc_mode=ALLOC_BUMP, lib_path=LIB_PATH):
        self.params = O.OrbParams(nfeatures, scale_factor, nlevels, ini_th, min_th, edge_th, im_w, im_h)
        self.nlevels = nlevels
        self.alloc_mode = alloc_mode
        self.L = lib(lib_path)
        self.cap = nfeatures + 3 * nlevels + 64 + 64 * nlevels
        self.last = {}

    def tables(self, w, h):
        nl = self.nlevels
        f = np.zeros(nl, np.int32); s = [np.zeros(nl, np.float32) for _ in range(4)]
        um = np.zeros(16, np.int32); edge = C.c_int(0); lw = np.zeros(nl, np.int32); lh = np.zeros(nl, np.int32)
        self.L.ref_orb_tables(C.byref(self.params), w, h, _p(f), _p(s[0]), _p(s[1]), _p(s[2]), _p(s[3]), _p(um), C.byref(edge), _p(lw), _p(lh))
        return dict(features_per_level=f, scale=s[0], inv_scale=s[1], sigma2=s[2], inv_sigma2=s[3], umax=um, edge=edge.value,
                    level_w=lw, level_h=lh)

    def extract(self, img, lapping=(0, 1000), want_desc=True, taps=False):
        """-> (ret, keypoints, descriptors or None); with taps=True self.last holds levels / blurred / FAST counters"""
        if img is None or img.size == 0:
            return -1, np.empty(0, O.KEYPOINT_DTYPE), None
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        kps = np.zeros(self.cap, O.KEYPOINT_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = C.c_int(0)
        t = np.zeros(8, np.int64)
        pyr = blur = present = None
        if taps:
            tb = self.tables(w, h)
            tot = int((tb["level_w"].astype(np.int64) * tb["level_h"]).sum())
            pyr = np.zeros(tot, np.uint8); blur = np.zeros(tot, np.uint8); present = np.zeros(self.nlevels, np.int32)
        ret = self.L.ref_orb_extract(C.byref(self.params), _p(img), w, h, img.strides[0], int(lapping[0]), int(lapping[1]), int(want_desc),
                                     self.alloc_mode, _p(kps), _p(desc), self.cap, C.byref(n), _p(pyr), _p(blur) if want_desc else None,
                                     _p(present), _p(t))
        assert n.value <= self.cap, "capacity exceeded"
        self.last = dict(fast_calls=int(t[0]), fast_calls_nonempty=int(t[1]), candidates=int(t[2]), bump_overflow=int(t[3]), arena_bytes=int(t[4]))
        assert self.last["bump_overflow"] == 0, "bump arena exhausted: pointer order no longer equals creation order"
        if taps:
            levels, blurred, o, ob = [], [], 0, 0
            for l in range(self.nlevels):
                lw, lh = int(tb["level_w"][l]), int(tb["level_h"][l])
                levels.append(pyr[o:o + lw * lh].reshape(lh, lw).copy()); o += lw * lh
                if want_desc and present[l]:
                    blurred.append(blur[ob:ob + lw * lh].reshape(lh, lw).copy()); ob += lw * lh
                else:
                    blurred.append(None)
            self.last.update(levels=levels, blurred=blurred)
        return ret, kps[:n.value].copy(), (desc[:n.value].copy() if want_desc else None)

    def tracked_desc(self, img, kps):
        img = np.ascontiguousarray(img, np.uint8)
        kps = np.ascontiguousarray(kps, O.KEYPOINT_DTYPE)
        desc = np.zeros((len(kps), 32), np.uint8)
        self.L.ref_orb_tracked_desc(C.byref(self.params), _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), len(kps), _p(desc))
        return desc

    def assign_level_by_best_desc(self, ref_desc, img, kps):
        img = np.ascontiguousarray(img, np.uint8)
        kps = np.ascontiguousarray(kps, O.KEYPOINT_DTYPE).copy()
        ref_desc = np.ascontiguousarray(ref_desc, np.uint8)
        self.L.ref_orb_assign_level_by_best_desc(C.byref(self.params), _p(ref_desc), _p(img), img.shape[1], img.shape[0], img.strides[0],
                                                 _p(kps), len(kps))
        return kps


def distribute_octtree(kx, ky, kresp, minX, maxX, minY, maxY, N, alloc_mode=ALLOC_BUMP) -> np.ndarray:
    kx = np.ascontiguousarray(kx, np.float32); ky = np.ascontiguousarray(ky, np.float32); kr = np.ascontiguousarray(kresp, np.float32)
    cap = len(kx) + 8
    out = np.empty(cap, np.int32)
    n = lib().ref_distribute_octtree(_p(kx), _p(ky), _p(kr), len(kx), minX, maxX, minY, maxY, N, alloc_mode, _p(out), cap)
    return out[:n].copy()


def descriptor_distance(a, b) -> int:
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().ref_descriptor_distance(_p(a), _p(b))


def descriptor_distance_matrix(q, db) -> np.ndarray:
    q = np.ascontiguousarray(q, np.uint8); db = np.ascontiguousarray(db, np.uint8)
    out = np.zeros((len(q), len(db)), np.int32)
    lib().ref_descriptor_distance_matrix(_p(q), len(q), _p(db), len(db), _p(out))
    return out


def orb_extract_batch_mt(frames, nthreads, want_desc=True, **kw):
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    p = O.OrbParams(kw.get("nfeatures", 1000), kw.get("scale_factor", 1.2), kw.get("nlevels", 8), kw.get("ini_th", 20),
                    kw.get("min_th", 7), kw.get("edge_th", 19), w, h)
    counts = np.zeros(n, np.int32)
    total = lib().ref_orb_extract_batch_mt(C.byref(p), _p(frames), n, w, h, nthreads, int(want_desc), _p(counts))
    return total, counts


# ----------------------------------------------------------------------------- events (EventConversion.cc, unmodified)
def ev_accumulate(evs, w, h, sigma=1.0, mode=1, Tcw=None, depth=1.0, K=None, se2=None, pol=False, normalize=False, kb8=None):
    """same contract as oracle_lib.ev_accumulate: -> (img_f32, (min, max) as the reference tracked them, u8 or None)"""
    evs = np.ascontiguousarray(evs)
    assert evs.dtype.itemsize == 24
    img = np.zeros((h, w), np.float32)
    u8 = np.zeros((h, w), np.uint8)
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16) if Tcw is not None else None
    Kc = np.ascontiguousarray(K, np.float32) if K is not None else None
    if kb8 is not None:
        Kc = np.ascontiguousarray(list(K) + list(kb8), np.float32)
    s2 = np.ascontiguousarray(se2, np.float32) if se2 is not None else None
    f = lib().ref_ev_accumulate_cam; f.restype = C.c_int
    r = f(_p(evs), C.c_int64(len(evs)), w, h, C.c_float(sigma), mode, _p(T), C.c_float(depth), _p(Kc), 0 if kb8 is None else 1, _p(s2),
          0 if s2 is None else len(s2), int(pol), int(normalize), _p(img), _p(u8))
    assert r >= 0
    return img, (u8 if r == 1 else None)


# ----------------------------------------------------------------------------- tracking-thread matchers (ORBmatcher.cc, cut out verbatim)
def search_by_projection(*a, **kw):
    return O.search_by_projection_ex(*a, L=lib(), fn="ref_search_by_projection", **kw)


def search_by_projection_reloc(*a, **kw):
    return O.search_by_projection_reloc(*a, L=lib(), fn="ref_search_by_projection_reloc", **kw)


def search_by_projection_map_points(*a, **kw):
    return O.search_by_projection_map_points_ex(*a, L=lib(), fn="ref_search_by_projection_map_points", **kw)


def search_for_initialization(kps1, desc1, kps2, desc2, bounds, prev_xy, window_size=100, nnratio=0.9, check_ori=True):
    k1 = np.ascontiguousarray(kps1, O.KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, O.KEYPOINT_DTYPE)
    d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32); pv = np.ascontiguousarray(prev_xy, np.float32).copy()
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    f = lib().ref_search_for_initialization; f.restype = C.c_int
    n = f(_p(k1), _p(d1), C.c_int(len(k1)), _p(k2), _p(d2), C.c_int(len(k2)), _p(b), _p(pv), C.c_int(window_size), C.c_float(nnratio),
          C.c_int(int(check_ori)), _p(m12))
    return n, m12[:len(k1)].copy(), pv


def search_by_bow(kps_kf, desc_kf, valid_kf, fv_kf, kps_f, desc_f, fv_f, nnratio=0.7, check_ori=True):
    k1 = np.ascontiguousarray(kps_kf, O.KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps_f, O.KEYPOINT_DTYPE)
    d1 = np.ascontiguousarray(desc_kf, np.uint8); d2 = np.ascontiguousarray(desc_f, np.uint8); v = np.ascontiguousarray(valid_kf, np.uint8)
    a = [np.ascontiguousarray(fv_kf[0], np.uint32), np.ascontiguousarray(fv_kf[1], np.int32), np.ascontiguousarray(fv_kf[2], np.uint32)]
    b = [np.ascontiguousarray(fv_f[0], np.uint32), np.ascontiguousarray(fv_f[1], np.int32), np.ascontiguousarray(fv_f[2], np.uint32)]
    mf = np.full(max(len(k2), 1), -1, np.int32)
    f = lib().ref_search_by_bow; f.restype = C.c_int
    n = f(_p(k1), _p(d1), _p(v), _p(a[0]), _p(a[1]), _p(a[2]), C.c_int(len(a[0])), C.c_int(len(k1)), _p(k2), _p(d2), C.c_int(len(k2)), _p(b[0]),
          _p(b[1]), _p(b[2]), C.c_int(len(b[0])), C.c_float(nnratio), C.c_int(int(check_ori)), _p(mf))
    return n, mf[:len(k2)].copy()


def search_by_bow_kf(*a, **kw):
    """the reference's own SearchByBoW(KeyFrame*, KeyFrame*, ...) body (cut into libref)"""
    return O.search_by_bow_kf(*a, _fn=lib().ref_search_by_bow_kf, **kw)


def features_in_area(kps, bounds, x, y, r, min_level=-1, max_level=-1):
    k = np.ascontiguousarray(kps, O.KEYPOINT_DTYPE); b = np.ascontiguousarray(bounds, np.float32)
    out = np.zeros(max(len(k), 1), np.int32)
    f = lib().ref_features_in_area; f.restype = C.c_int
    n = f(_p(k), C.c_int(len(k)), _p(b), C.c_float(x), C.c_float(y), C.c_float(r), C.c_int(min_level), C.c_int(max_level), _p(out), C.c_int(len(out)))
    return out[:n].copy()


def lk_refine(curr_xy, status, ref_kps, w, h, first_octave_only=False, matches12=None, cnt_matches=None):
    """the reference's own ELK_Tracker::refineTrackedPts (+ refineFirstOctaveLevel) (KLT_Tracker.cpp:105-183, cut into libref)
    -> (nMatches, tracked, matches12, cnt_matches, px_disp); matches12 / cnt_matches given = the caller's non-empty vectors"""
    f = lib().ref_lk_refine; f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p] * 5
    cur = np.ascontiguousarray(curr_xy, np.float32).reshape(-1, 2); st = np.ascontiguousarray(status, np.uint8)
    rk = np.ascontiguousarray(ref_kps, O.KEYPOINT_DTYPE); n = len(rk)
    have = matches12 is not None
    m12 = np.ascontiguousarray(matches12, np.int32).copy() if have else np.zeros(n, np.int32)
    cnt = np.ascontiguousarray(cnt_matches, np.int32).copy() if have else np.zeros(n, np.int32)
    tr = np.zeros(n, O.KEYPOINT_DTYPE); disp = np.zeros(n, np.float32); c2 = np.zeros(2, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = f(p(cur), p(st), p(rk), n, w, h, 1 if first_octave_only else 0, 1 if have else 0, p(tr), p(m12), p(cnt), p(disp), p(c2))
    assert rc == 0
    return int(c2[0]), tr, m12, cnt, disp[:c2[1]].copy()


# ----------------------------------------------------------------------------- keyframe-side searches (oracle/ref_guided_kf_api.cc)
def _kf_common(c):
    k2 = np.ascontiguousarray(c["kps2"], O.KEYPOINT_DTYPE); d2 = np.ascontiguousarray(c["desc2"], np.uint8)
    b = np.ascontiguousarray(c["bounds"], np.float32); K = np.ascontiguousarray(c["K"], np.float32); sf = np.ascontiguousarray(c["scale_factors"], np.float32)
    return k2, d2, b, K, sf


def search_by_projection_kf(c, overload):
    """ORBmatcher::SearchByProjection(KeyFrame*, Scw, ...) (:480 / :595) on a case of tests/kf_cases.py -> (nmatches, match2[n2])"""
    k2, d2, b, K, sf = _kf_common(c)
    pt = np.ascontiguousarray(c["pt8"], np.float32); lv = np.ascontiguousarray(c["level1"], np.int32); fl = np.ascontiguousarray(c["flags1"], np.uint8)
    dm = np.ascontiguousarray(c["descMP"], np.uint8); hd = np.ascontiguousarray(c["held_id2"], np.int32)
    m2 = np.full(max(len(k2), 1), -1, np.int32)
    f = lib().ref_search_by_projection_kf; f.restype = C.c_int
    n = f(C.c_int(overload), _p(pt), _p(lv), _p(fl), _p(dm), C.c_int(len(lv)), _p(k2), _p(d2), _p(hd), C.c_int(len(k2)), _p(b), _p(K), _p(sf),
          C.c_int(len(sf)), C.c_int(int(c["th"])), C.c_float(c["ratio"]), _p(m2))
    return n, m2[:len(k2)].copy()


def fuse(c, overload):
    """ORBmatcher::Fuse (:1407 / :1619) -> (nFused, events[k, 3])"""
    k2, d2, b, K, sf = _kf_common(c)
    pt = np.ascontiguousarray(c["pt8"], np.float32); lv = np.ascontiguousarray(c["level1"], np.int32); fl = np.ascontiguousarray(c["flags1"], np.uint8)
    ob = np.ascontiguousarray(c["obs1"], np.int32); pr = np.ascontiguousarray(c["present1"], np.uint8); dm = np.ascontiguousarray(c["descMP"], np.uint8)
    oc = np.ascontiguousarray(c["occupied2"], np.uint8); oo = np.ascontiguousarray(c["occ_obs2"], np.int32)
    ur = None if c.get("u_right2") is None else np.ascontiguousarray(c["u_right2"], np.float32)
    inv = np.ascontiguousarray(c["inv_sigma2"], np.float32)
    cap = len(lv) + 8
    ev = np.zeros((cap, 3), np.int32); ne = C.c_int(0)
    f = lib().ref_fuse; f.restype = C.c_int
    n = f(C.c_int(overload), _p(pt), _p(lv), _p(ob), _p(fl), _p(pr), _p(dm), C.c_int(len(lv)), _p(k2), _p(d2), _p(oc), _p(oo), _p(ur), C.c_int(len(k2)),
          _p(b), _p(K), C.c_float(c["mbf"]), _p(sf), _p(inv), C.c_int(len(sf)), C.c_float(c["th"]), _p(ev), C.c_int(cap), C.byref(ne))
    assert ne.value <= cap
    return n, ev[:ne.value].copy()


def search_by_sim3(c):
    """ORBmatcher::SearchBySim3 (:1743) -> (nFound, match12[n1])"""
    k2, d2, b, K, sf = _kf_common(c)
    k1 = np.ascontiguousarray(c["kps1"], O.KEYPOINT_DTYPE); d1 = np.ascontiguousarray(c["desc1"], np.uint8)
    a = [np.ascontiguousarray(c[k + "_1"], t) for k, t in (("pt8", np.float32), ("level", np.int32), ("flags", np.uint8), ("present", np.uint8))]
    e = [np.ascontiguousarray(c[k + "_2"], t) for k, t in (("pt8", np.float32), ("level", np.int32), ("flags", np.uint8), ("present", np.uint8))]
    mi = np.ascontiguousarray(c["matched12_in"], np.int32); m12 = np.full(max(len(k1), 1), -1, np.int32)
    f = lib().ref_search_by_sim3; f.restype = C.c_int
    n = f(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(k1), _p(d1), C.c_int(len(k1)), _p(e[0]), _p(e[1]), _p(e[2]), _p(e[3]), _p(k2), _p(d2),
          C.c_int(len(k2)), _p(mi), _p(b), _p(K), _p(sf), C.c_int(len(sf)), C.c_float(c["th"]), _p(m12))
    return n, m12[:len(k1)].copy()


def search_for_triangulation(kps1, desc1, has_mp1, u_right1, fv1, kps2, desc2, has_mp2, u_right2, fv2, K4, t1w, t2w, scale_factors, level_sigma2,
                             only_stereo=False, coarse=False, check_ori=True):
    """ORBmatcher::SearchForTriangulation (:975-1214) -> (nmatches, match12[n1], F12[9], epipole[2])"""
    k1 = np.ascontiguousarray(kps1, O.KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, O.KEYPOINT_DTYPE)
    d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    h1 = np.ascontiguousarray(has_mp1, np.uint8); h2 = np.ascontiguousarray(has_mp2, np.uint8)
    u1 = None if u_right1 is None else np.ascontiguousarray(u_right1, np.float32); u2 = None if u_right2 is None else np.ascontiguousarray(u_right2, np.float32)
    a, b = O._fv(fv1), O._fv(fv2)
    K = np.ascontiguousarray(K4, np.float32); t1 = np.ascontiguousarray(t1w, np.float32); t2 = np.ascontiguousarray(t2w, np.float32)
    sc = np.ascontiguousarray(scale_factors, np.float32); sg = np.ascontiguousarray(level_sigma2, np.float32)
    m12 = np.full(max(len(k1), 1), -1, np.int32); F = np.zeros(9, np.float32); ep = np.zeros(2, np.float32)
    f = lib().ref_search_for_triangulation; f.restype = C.c_int
    n = f(_p(k1), _p(d1), _p(h1), _p(u1), C.c_int(len(k1)), _p(a[0]), _p(a[1]), _p(a[2]), C.c_int(len(a[0])), _p(k2), _p(d2), _p(h2), _p(u2), C.c_int(len(k2)),
          _p(b[0]), _p(b[1]), _p(b[2]), C.c_int(len(b[0])), _p(K), _p(t1), _p(t2), _p(sc), _p(sg), C.c_int(len(sc)), C.c_int(int(only_stereo)),
          C.c_int(int(coarse)), C.c_int(int(check_ori)), _p(m12), _p(F), _p(ep))
    return n, m12[:len(k1)].copy(), F, ep

