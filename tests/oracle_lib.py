"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg — never by
the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")]
)
MATCH_DTYPE = np.dtype([("best_dist", "<i4"), ("best_idx", "<i4"), ("second_dist", "<i4"), ("accepted", "<i4")])


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scaleFactor", C.c_float), ("nlevels", C.c_int), ("iniThFAST", C.c_int),
                ("minThFAST", C.c_int), ("edgeTh", C.c_int), ("imW", C.c_int), ("imH", C.c_int)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("orb_oracle.cc", "match_oracle.cc", "guided_oracle.cc", "bow_oracle.cc", "event_oracle.cc", "lk_oracle.cc", "oracle.h",
                                                  "brief_pattern_31.inc", "Makefile")]
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    u8p, i32p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_float)
    vp = C.c_void_p
    L.orc_cv_round_f.argtypes = [C.c_float]; L.orc_cv_round_f.restype = C.c_int
    L.orc_resize_linear_u8.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, C.c_int, C.c_size_t]
    L.orc_resize_linear_u8.restype = None
    L.orc_copy_make_border_reflect101.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, C.c_size_t]
    L.orc_copy_make_border_reflect101.restype = None
    L.orc_gauss5x5_s2_u8.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_size_t]; L.orc_gauss5x5_s2_u8.restype = None
    L.orc_fast9_16.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, vp, vp, vp, C.c_int]
    L.orc_fast9_16.restype = C.c_int
    L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]; L.orc_fast_atan2.restype = C.c_float
    L.orc_distribute_octtree.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    L.orc_distribute_octtree.restype = C.c_int
    L.orc_orb_create.argtypes = [C.POINTER(OrbParams)]; L.orc_orb_create.restype = vp
    L.orc_orb_destroy.argtypes = [vp]; L.orc_orb_destroy.restype = None
    L.orc_orb_edge_threshold.argtypes = [vp]; L.orc_orb_edge_threshold.restype = C.c_int
    L.orc_orb_features_per_level.argtypes = [vp, vp]; L.orc_orb_features_per_level.restype = C.c_int
    L.orc_orb_scale_factors.argtypes = [vp, vp, vp, vp, vp]; L.orc_orb_scale_factors.restype = C.c_int
    L.orc_orb_umax.argtypes = [vp, vp]; L.orc_orb_umax.restype = C.c_int
    L.orc_orb_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp]
    L.orc_orb_extract.restype = C.c_int
    L.orc_orb_level_size.argtypes = [vp, C.c_int, vp, vp]; L.orc_orb_level_size.restype = C.c_int
    L.orc_orb_get_level.argtypes = [vp, C.c_int, vp, C.c_size_t]; L.orc_orb_get_level.restype = C.c_int
    L.orc_orb_get_blurred.argtypes = [vp, C.c_int, vp, C.c_size_t]; L.orc_orb_get_blurred.restype = C.c_int
    L.orc_orb_num_candidates.argtypes = [vp, C.c_int]; L.orc_orb_num_candidates.restype = C.c_int
    L.orc_orb_get_candidates.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int]; L.orc_orb_get_candidates.restype = C.c_int
    L.orc_orb_num_level_kps.argtypes = [vp, C.c_int]; L.orc_orb_num_level_kps.restype = C.c_int
    L.orc_orb_get_level_kps.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int]; L.orc_orb_get_level_kps.restype = C.c_int
    L.orc_orb_num_fallback_cells.argtypes = [vp]; L.orc_orb_num_fallback_cells.restype = C.c_int
    L.orc_orb_extract_batch_mt.argtypes = [C.POINTER(OrbParams), vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.orc_orb_extract_batch_mt.restype = C.c_long
    L.orc_orb_tracked_desc.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, vp]; L.orc_orb_tracked_desc.restype = C.c_int
    L.orc_orb_assign_level_by_best_desc.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int]
    L.orc_orb_assign_level_by_best_desc.restype = C.c_int
    L.orc_descriptor_distance.argtypes = [vp, vp]; L.orc_descriptor_distance.restype = C.c_int
    L.orc_hamming_best2.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int, C.c_float, C.c_int, vp, C.c_int]
    L.orc_hamming_best2.restype = None
    L.orc_rotation_filter.argtypes = [vp, vp, vp, C.c_int]; L.orc_rotation_filter.restype = C.c_int
    L.orc_ev_accumulate.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, vp, C.c_float, vp, vp, C.c_int,
                                    C.c_int, C.c_int, vp, vp]
    L.orc_ev_accumulate.restype = C.c_int
    L.orc_normalize_convert_u8.argtypes = [vp, C.c_int, C.c_float, C.c_float, vp]; L.orc_normalize_convert_u8.restype = None
    L.orc_normalize_minmax_u8.argtypes = [vp, C.c_int, vp]; L.orc_normalize_minmax_u8.restype = None
    L.orc_ev_mci_jac.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_float, vp, C.c_float, vp, C.c_int, C.c_int, vp]; L.orc_ev_mci_jac.restype = None
    L.orc_image_focus.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]; L.orc_image_focus.restype = C.c_float
    L.orc_pyrdown_u8.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, C.c_int, C.c_size_t]; L.orc_pyrdown_u8.restype = None
    L.orc_scharr_deriv.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, vp]; L.orc_scharr_deriv.restype = None
    L.orc_lk_track.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                               C.c_float, vp, vp]
    L.orc_lk_track.restype = C.c_int
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------- primitives
def resize_linear(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dst.strides[0])
    return dst


def border_reflect101(src: np.ndarray, b: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((src.shape[0] + 2 * b, src.shape[1] + 2 * b), np.uint8)
    lib().orc_copy_make_border_reflect101(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), b, dst.strides[0])
    return dst


def gauss5(src: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_gauss5x5_s2_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast(img: np.ndarray, threshold: int, nms: bool = True):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    xs = np.empty(cap, np.int32); ys = np.empty(cap, np.int32); sc = np.empty(cap, np.int32)
    n = lib().orc_fast9_16(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, int(nms), _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def fast_atan2(y: float, x: float) -> float:
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


def distribute_octtree(kx, ky, kresp, minX, maxX, minY, maxY, N) -> np.ndarray:
    kx = np.ascontiguousarray(kx, np.float32); ky = np.ascontiguousarray(ky, np.float32)
    kr = np.ascontiguousarray(kresp, np.float32)
    cap = len(kx) + 8
    out = np.empty(cap, np.int32)
    n = lib().orc_distribute_octtree(_p(kx), _p(ky), _p(kr), len(kx), minX, maxX, minY, maxY, N, _p(out), cap)
    return out[:n].copy()


# ----------------------------------------------------------------------------- extractor
class OrbOracle:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, edge_th=19, im_w=752, im_h=480):
        self.params = OrbParams(nfeatures, scale_factor, nlevels, ini_th, min_th, edge_th, im_w, im_h)
        self.h = lib().orc_orb_create(C.byref(self.params))
        assert self.h
        self.nlevels = nlevels
        self.cap = nfeatures + 3 * nlevels + 64 + 64 * nlevels   # elongated frames: up to 4 * round(aspect) nodes per level when the quota is tiny

    def __del__(self):
        try:
            if self.h:
                lib().orc_orb_destroy(self.h); self.h = None
        except Exception:
            pass

    @property
    def edge(self):
        return lib().orc_orb_edge_threshold(self.h)

    def features_per_level(self):
        out = np.empty(self.nlevels, np.int32)
        lib().orc_orb_features_per_level(self.h, _p(out))
        return out

    def scale_factors(self):
        a = [np.empty(self.nlevels, np.float32) for _ in range(4)]
        lib().orc_orb_scale_factors(self.h, *[_p(x) for x in a])
        return a

    def umax(self):
        out = np.empty(16, np.int32)
        lib().orc_orb_umax(self.h, _p(out))
        return out

    def extract(self, img: np.ndarray, lapping=(0, 1000), want_desc=True):
        """-> (ret, keypoints[KEYPOINT_DTYPE], descriptors (n,32) u8 or None)"""
        if img is None or img.size == 0:
            return -1, np.empty(0, KEYPOINT_DTYPE), None
        img = np.ascontiguousarray(img, np.uint8)
        kps = np.zeros(self.cap, KEYPOINT_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = C.c_int(0)
        ret = lib().orc_orb_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], int(lapping[0]),
                                    int(lapping[1]), int(want_desc), _p(kps), _p(desc), self.cap, C.byref(n))
        assert ret != -2, "oracle capacity exceeded"
        return ret, kps[:n.value].copy(), (desc[:n.value].copy() if want_desc else None)

    def level_size(self, l):
        w = C.c_int(); h = C.c_int()
        lib().orc_orb_level_size(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level(self, l):
        w, h = self.level_size(l)
        out = np.empty((h, w), np.uint8)
        lib().orc_orb_get_level(self.h, l, _p(out), out.strides[0])
        return out

    def blurred(self, l):
        w, h = self.level_size(l)
        out = np.empty((h, w), np.uint8)
        r = lib().orc_orb_get_blurred(self.h, l, _p(out), out.strides[0])
        return out if r == 0 else None

    def candidates(self, l):
        n = lib().orc_orb_num_candidates(self.h, l)
        xs = np.empty(n, np.int32); ys = np.empty(n, np.int32); sc = np.empty(n, np.int32)
        lib().orc_orb_get_candidates(self.h, l, _p(xs), _p(ys), _p(sc), n)
        return xs, ys, sc

    def level_kps(self, l):
        n = lib().orc_orb_num_level_kps(self.h, l)
        xs = np.empty(n, np.int32); ys = np.empty(n, np.int32); sc = np.empty(n, np.int32); an = np.empty(n, np.float32)
        lib().orc_orb_get_level_kps(self.h, l, _p(xs), _p(ys), _p(sc), _p(an), n)
        return xs, ys, sc, an

    def fallback_cells(self):
        return lib().orc_orb_num_fallback_cells(self.h)

    def tracked_desc(self, img, kps):
        img = np.ascontiguousarray(img, np.uint8)
        kps = np.ascontiguousarray(kps, KEYPOINT_DTYPE)
        desc = np.zeros((len(kps), 32), np.uint8)
        lib().orc_orb_tracked_desc(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), len(kps), _p(desc))
        return desc

    def assign_level_by_best_desc(self, ref_desc, img, kps):
        img = np.ascontiguousarray(img, np.uint8)
        kps = np.ascontiguousarray(kps, KEYPOINT_DTYPE).copy()
        ref_desc = np.ascontiguousarray(ref_desc, np.uint8)
        lib().orc_orb_assign_level_by_best_desc(self.h, _p(ref_desc), _p(img), img.shape[1], img.shape[0], img.strides[0],
                                                _p(kps), len(kps))
        return kps


def orb_extract_batch_mt(frames: np.ndarray, nthreads: int, want_desc=True, **kw):
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    p = OrbParams(kw.get("nfeatures", 1000), kw.get("scale_factor", 1.2), kw.get("nlevels", 8), kw.get("ini_th", 20),
                  kw.get("min_th", 7), kw.get("edge_th", 19), w, h)
    counts = np.zeros(n, np.int32)
    total = lib().orc_orb_extract_batch_mt(C.byref(p), _p(frames), n, w, h, nthreads, int(want_desc), _p(counts))
    return total, counts


# ----------------------------------------------------------------------------- matcher
def descriptor_distance(a, b) -> int:
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_descriptor_distance(_p(a), _p(b))


def hamming_best2(q, db, th=50, ratio=0.7, ratio_mode=0, nthreads=8):
    q = np.ascontiguousarray(q, np.uint8); db = np.ascontiguousarray(db, np.uint8)
    out = np.zeros(len(q), MATCH_DTYPE)
    lib().orc_hamming_best2(_p(q), len(q), _p(db), len(db), th, ratio, ratio_mode, _p(out), nthreads)
    return out


def rotation_filter(angle1, angle2, match12):
    a1 = np.ascontiguousarray(angle1, np.float32); a2 = np.ascontiguousarray(angle2, np.float32)
    m = np.ascontiguousarray(match12, np.int32).copy()
    n = lib().orc_rotation_filter(_p(a1), _p(a2), _p(m), len(m))
    return n, m


# ----------------------------------------------------------------------------- events
def ev_accumulate(evs, w, h, sigma=1.0, mode=1, Tcw=None, depth=1.0, K=None, se2=None, pol=False, normalize=False, kb8=None):
    """-> (img_f32 (h,w), (min,max), u8 or None).  kb8 = (k1, k2, k3, k4): the camera is a KannalaBrandt8 with K = fx, fy, cx, cy"""
    evs = np.ascontiguousarray(evs)
    assert evs.dtype.itemsize == 24
    img = np.zeros((h, w), np.float32)
    mm = np.zeros(2, np.float32)
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16) if Tcw is not None else None
    Kc = np.ascontiguousarray(K, np.float32) if K is not None else None
    if kb8 is not None:
        Kc = np.ascontiguousarray(list(K) + list(kb8), np.float32)
    s2 = np.ascontiguousarray(se2, np.float32) if se2 is not None else None
    f = lib().orc_ev_accumulate_cam; f.restype = C.c_int
    r = f(_p(evs), C.c_int64(len(evs)), w, h, C.c_float(sigma), mode, _p(T), C.c_float(depth), _p(Kc), 0 if kb8 is None else 1, _p(s2),
          0 if s2 is None else len(s2), int(pol), int(normalize), _p(img), _p(mm))
    assert r >= 0
    u8 = None
    if r == 1:
        u8 = np.empty((h, w), np.uint8)
        lib().orc_normalize_convert_u8(_p(img), w * h, float(mm[1]), float(mm[0]), _p(u8))
    return img, (float(mm[0]), float(mm[1])), u8


def normalize_minmax_u8(img):
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty(img.shape, np.uint8)
    lib().orc_normalize_minmax_u8(_p(img), img.size, _p(out))
    return out


# ----------------------------------------------------------------------------- pyramidal LK
def pyrdown(src: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyrdown_u8(_p(src), w, h, src.strides[0], _p(dst), dst.shape[1], dst.shape[0], dst.strides[0])
    return dst


def scharr_deriv(src: np.ndarray) -> np.ndarray:
    """(h, w, 2) int16: [..., 0] = Ix, [..., 1] = Iy (calcSharrDeriv)"""
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.zeros((h, w, 2), np.int16)
    lib().orc_scharr_deriv(_p(src), w, h, src.strides[0], _p(dst))
    return dst


def lk_track(prev_img, next_img, prev_pts, next_pts=None, win=23, max_level=1, max_iter=10, eps=0.03, min_eig=1e-4):
    """cv::calcOpticalFlowPyrLK restatement -> (next_pts float32[n,2], status uint8[n], err float32[n], levels_used)"""
    prev_img = np.ascontiguousarray(prev_img, np.uint8); next_img = np.ascontiguousarray(next_img, np.uint8)
    assert prev_img.shape == next_img.shape
    h, w = prev_img.shape
    pp = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    n = len(pp)
    use_init = next_pts is not None
    npts = np.ascontiguousarray(next_pts, np.float32).reshape(-1, 2).copy() if use_init else np.zeros((n, 2), np.float32)
    status = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
    lv = lib().orc_lk_track(_p(prev_img), _p(next_img), w, h, prev_img.strides[0], _p(pp), _p(npts), n, win, max_level, max_iter, float(eps),
                            1 if use_init else 0, float(min_eig), _p(status), _p(err))
    return npts, status, err, lv


def lk_refine(curr_xy, status, ref_kps, w, h, first_octave_only=False, matches12=None, cnt_matches=None):
    """ELK_Tracker::refineTrackedPts (+ refineFirstOctaveLevel) restatement (KLT_Tracker.cpp:105-183)
    -> (nMatches, tracked, matches12, cnt_matches, px_disp).  Without matches12 / cnt_matches the caller's vectors are empty, i.e.
    they start as -1 / 1 (resize(n, -1) / resize(n, 1), :113-133)."""
    f = lib().orc_lk_refine; f.restype = None
    f.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p] * 5
    cur = np.ascontiguousarray(curr_xy, np.float32).reshape(-1, 2); st = np.ascontiguousarray(status, np.uint8)
    rk = np.ascontiguousarray(ref_kps, KEYPOINT_DTYPE); n = len(rk)
    m12 = np.ascontiguousarray(matches12, np.int32).copy() if matches12 is not None else np.full(n, -1, np.int32)
    cnt = np.ascontiguousarray(cnt_matches, np.int32).copy() if cnt_matches is not None else np.ones(n, np.int32)
    tr = np.zeros(n, KEYPOINT_DTYPE); disp = np.zeros(n, np.float32); c2 = np.zeros(2, np.int32)
    f(_p(cur), _p(st), _p(rk), n, w, h, 1 if first_octave_only else 0, _p(tr), _p(m12), _p(cnt), _p(disp), _p(c2))
    return int(c2[0]), tr, m12, cnt, disp[:c2[1]].copy()


# ----------------------------------------------------------------------------- contrast metric
FOCUS_LOCAL_STD, FOCUS_GLOBAL_STD, FOCUS_LOCAL_MEAN = 0, 1, 2


def image_focus(img: np.ndarray, what=FOCUS_LOCAL_STD, avg=True, patch=30) -> float:
    img = np.ascontiguousarray(img, np.float32)
    return float(lib().orc_image_focus(_p(img), img.shape[1], img.shape[0], patch, what, 1 if avg else 0))


def ev_mci_jac(evs, w, h, sigma, R, t, med_depth, K, pol=False, global_mean=False) -> np.ndarray:
    """ev2mci_gg_f_jac -> float64[6] (d contrast / d [wx wy wz vx vy vz])"""
    evs = np.ascontiguousarray(evs)
    Rt = np.concatenate([np.asarray(R, np.float64).reshape(9), np.asarray(t, np.float64).reshape(3)])
    Kc = np.ascontiguousarray(K, np.float32)
    out = np.zeros(6, np.float64)
    lib().orc_ev_mci_jac(_p(evs), len(evs), w, h, float(sigma), _p(Rt), float(med_depth), _p(Kc), int(pol), int(global_mean), _p(out))
    return out


# ---- guided matching
GRID_COLS, GRID_ROWS = 64, 48


def frame_grid(kps, bounds):
    """Frame::AssignFeaturesToGrid -> (cell_start[3073], cell_idx[assigned])"""
    L = lib()
    kps = np.ascontiguousarray(kps, KEYPOINT_DTYPE); b = np.ascontiguousarray(bounds, np.float32)
    cs = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32); ci = np.zeros(max(len(kps), 1), np.int32)
    L.orc_frame_grid.restype = C.c_int
    n = L.orc_frame_grid(_p(kps), C.c_int(len(kps)), _p(b), _p(cs), _p(ci))
    return cs, ci[:n].copy()


def features_in_area(kps, bounds, cs, ci, x, y, r, min_level=0, max_level=-1):
    L = lib()
    kps = np.ascontiguousarray(kps, KEYPOINT_DTYPE); b = np.ascontiguousarray(bounds, np.float32)
    ci = np.ascontiguousarray(ci, np.int32) if len(ci) else np.zeros(1, np.int32)
    out = np.zeros(max(len(kps), 1), np.int32)
    L.orc_features_in_area.restype = C.c_int
    n = L.orc_features_in_area(_p(kps), C.c_int(len(kps)), _p(b), _p(cs), _p(ci), C.c_float(x), C.c_float(y), C.c_float(r),
                               C.c_int(min_level), C.c_int(max_level), _p(out), C.c_int(len(out)))
    return out[:n].copy()


def search_for_initialization(kps1, desc1, kps2, desc2, bounds, prev_xy, window_size=100, nnratio=0.9, check_ori=True):
    """ORBmatcher::SearchForInitialization -> (nmatches, matches12[n1], prev_xy updated copy)"""
    L = lib()
    kps1 = np.ascontiguousarray(kps1, KEYPOINT_DTYPE); kps2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
    d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32)
    prev = np.array(prev_xy, np.float32, copy=True).reshape(-1, 2)
    m12 = np.full(max(len(kps1), 1), -1, np.int32)
    L.orc_search_for_initialization.restype = C.c_int
    n = L.orc_search_for_initialization(_p(kps1), _p(d1), C.c_int(len(kps1)), _p(kps2), _p(d2), C.c_int(len(kps2)), _p(b), _p(prev),
                                        C.c_int(window_size), C.c_float(nnratio), C.c_int(int(check_ori)), _p(m12))
    return n, m12[:len(kps1)].copy(), prev


def search_by_projection(x3Dc, valid1, obs1, kps1, descMP, kps2, desc2, bounds, K4, scale_factors, th=15.0, check_ori=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono=True) -> (nmatches, match_cur[n2])"""
    L = lib()
    x = np.ascontiguousarray(x3Dc, np.float32).reshape(-1, 3); v = np.ascontiguousarray(valid1, np.uint8); o = np.ascontiguousarray(obs1, np.int32)
    k1 = np.ascontiguousarray(kps1, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
    dm = np.ascontiguousarray(descMP, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32); K = np.ascontiguousarray(K4, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    mc = np.full(max(len(k2), 1), -1, np.int32)
    L.orc_search_by_projection.restype = C.c_int
    n = L.orc_search_by_projection(_p(x), _p(v), _p(o), _p(k1), _p(dm), C.c_int(len(k1)), _p(k2), _p(d2), C.c_int(len(k2)), _p(b), _p(K),
                                   _p(sf), C.c_int(len(sf)), C.c_float(th), C.c_int(int(check_ori)), _p(mc))
    return n, mc[:len(k2)].copy()


def _guided_args(kps1, kps2, descMP, desc2, bounds, scale_factors):
    k1 = np.ascontiguousarray(kps1, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
    dm = np.ascontiguousarray(descMP, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    return k1, k2, dm, d2, b, sf


def search_by_projection_ex(x3Dc, valid1, obs1, kps1, descMP, kps2, desc2, bounds, K4, scale_factors, th=15.0, check_ori=True, level_mode=0,
                            mbf=0.0, u_right2=None, L=None, fn="orc_search_by_projection_ex"):
    """SearchByProjection(CurrentFrame, LastFrame, th, bMono) with the rectified-stereo branches -> (nmatches, match_cur[n2]).
    L / fn select the library (the oracle by default; tests/ref_lib.py passes libref and the ref_ name)"""
    L = L or lib()
    x = np.ascontiguousarray(x3Dc, np.float32).reshape(-1, 3); v = np.ascontiguousarray(valid1, np.uint8); o = np.ascontiguousarray(obs1, np.int32)
    k1, k2, dm, d2, b, sf = _guided_args(kps1, kps2, descMP, desc2, bounds, scale_factors)
    K = np.ascontiguousarray(K4, np.float32)
    ur = None if u_right2 is None else np.ascontiguousarray(u_right2, np.float32)
    mc = np.full(max(len(k2), 1), -1, np.int32)
    f = getattr(L, fn); f.restype = C.c_int
    n = f(_p(x), _p(v), _p(o), _p(k1), _p(dm), C.c_int(len(k1)), _p(k2), _p(d2), C.c_int(len(k2)), _p(b), _p(K), _p(sf), C.c_int(len(sf)),
          C.c_float(th), C.c_int(int(check_ori)), C.c_int(level_mode), C.c_float(mbf), _p(ur), _p(mc))
    return n, mc[:len(k2)].copy()


def search_by_projection_reloc(x3Dc, valid1, level1, kps1, descMP, kps2, desc2, held2, bounds, K4, scale_factors, th=10.0, orb_dist=100,
                               check_ori=True, L=None, fn="orc_search_by_projection_reloc"):
    """SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (Tracking::Relocalization) -> (nmatches, match_cur[n2])"""
    L = L or lib()
    x = np.ascontiguousarray(x3Dc, np.float32).reshape(-1, 3); v = np.ascontiguousarray(valid1, np.uint8); lv = np.ascontiguousarray(level1, np.int32)
    k1, k2, dm, d2, b, sf = _guided_args(kps1, kps2, descMP, desc2, bounds, scale_factors)
    K = np.ascontiguousarray(K4, np.float32)
    hd = None if held2 is None else np.ascontiguousarray(held2, np.uint8)
    mc = np.full(max(len(k2), 1), -1, np.int32)
    f = getattr(L, fn); f.restype = C.c_int
    n = f(_p(x), _p(v), _p(lv), _p(k1), _p(dm), C.c_int(len(k1)), _p(k2), _p(d2), _p(hd), C.c_int(len(k2)), _p(b), _p(K), _p(sf), C.c_int(len(sf)),
          C.c_float(th), C.c_int(int(orb_dist)), C.c_int(int(check_ori)), _p(mc))
    return n, mc[:len(k2)].copy()


def _fv(fv):
    return [np.ascontiguousarray(fv[0], np.uint32), np.ascontiguousarray(fv[1], np.int32), np.ascontiguousarray(fv[2], np.uint32)]


def triangulation_flags(has_mp, u_right, only_stereo):
    """per-feature flags of SearchForTriangulation: bit 0 = takes part (no map point; stereo when bOnlyStereo), bit 1 = bStereo"""
    has_mp = np.asarray(has_mp).astype(bool)
    st = np.zeros(len(has_mp), bool) if u_right is None else (np.asarray(u_right, np.float32) >= 0)
    part = ~has_mp & (st if only_stereo else True)
    return (part.astype(np.uint8) | (st.astype(np.uint8) << 1)).astype(np.uint8)


def search_for_triangulation(kps1, desc1, flags1, fv1, kps2, desc2, flags2, fv2, F12, ep, scale2, sigma2, coarse=False, check_ori=True):
    """ORBmatcher::SearchForTriangulation (pinhole keyframes) -> (nmatches, match12[n1])"""
    k1 = np.ascontiguousarray(kps1, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
    d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    f1 = np.ascontiguousarray(flags1, np.uint8); f2 = np.ascontiguousarray(flags2, np.uint8)
    a, b = _fv(fv1), _fv(fv2)
    F = np.ascontiguousarray(F12, np.float32).reshape(9); e = np.ascontiguousarray(ep, np.float32)
    sc = np.ascontiguousarray(scale2, np.float32); sg = np.ascontiguousarray(sigma2, np.float32)
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    f = lib().orc_search_for_triangulation; f.restype = C.c_int
    n = f(_p(k1), _p(d1), _p(f1), C.c_int(len(k1)), _p(a[0]), _p(a[1]), _p(a[2]), C.c_int(len(a[0])), _p(k2), _p(d2), _p(f2), C.c_int(len(k2)),
          _p(b[0]), _p(b[1]), _p(b[2]), C.c_int(len(b[0])), _p(F), _p(e), _p(sc), _p(sg), C.c_int(len(sc)), C.c_int(int(coarse)), C.c_int(int(check_ori)), _p(m12))
    return n, m12[:len(k1)].copy()


AREA_QUERY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("r", "<f4"), ("min_level", "<i4"), ("max_level", "<i4")])


def search_windows(queries, ur, descMP, kps2, desc2, held2, u_right2, bounds, query_min_xy=None, inv_level_sigma2=None, blocking=False,
                   th_high=50):
    """matching core of the keyframe-side searches (orc_search_windows) -> (nmatches, best_idx[n1], best_dist[n1], match2[n2])"""
    q = np.ascontiguousarray(queries, AREA_QUERY_DTYPE)
    k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE); dm = np.ascontiguousarray(descMP, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32)
    opt = lambda a, t: None if a is None else np.ascontiguousarray(a, t)
    urq, hd, ur2, qm, inv = opt(ur, np.float32), opt(held2, np.uint8), opt(u_right2, np.float32), opt(query_min_xy, np.float32), opt(inv_level_sigma2, np.float32)
    bi = np.full(max(len(q), 1), -1, np.int32); bd = np.full(max(len(q), 1), 256, np.int32); m2 = np.full(max(len(k2), 1), -1, np.int32)
    f = lib().orc_search_windows; f.restype = C.c_int
    n = f(_p(q), _p(urq), _p(dm), C.c_int(len(q)), _p(k2), _p(d2), _p(hd), _p(ur2), C.c_int(len(k2)), _p(b), _p(qm), _p(inv),
          C.c_int(0 if inv is None else len(inv)), C.c_int(int(blocking)), C.c_int(int(th_high)), _p(bi), _p(bd), _p(m2))
    return n, bi[:len(q)].copy(), bd[:len(q)].copy(), m2[:len(k2)].copy()


def search_by_projection_map_points_ex(pts, proj_xr, descMP, kps2, desc2, held2, u_right2, bounds, scale_factors, th=1.0, far_points=False,
                                       th_far=0.0, nnratio=0.8, L=None, fn="orc_search_by_projection_map_points_ex"):
    """SearchByProjection(F, vpMapPoints, ...) with the rectified-stereo test -> (nmatches, match_cur[n2])"""
    from eorb_slam_b200.synth import TRACK_POINT_DTYPE
    L = L or lib()
    p = np.ascontiguousarray(pts, TRACK_POINT_DTYPE)
    _, k2, dm, d2, b, sf = _guided_args(np.zeros(0, KEYPOINT_DTYPE), kps2, descMP, desc2, bounds, scale_factors)
    hd = None if held2 is None else np.ascontiguousarray(held2, np.uint8)
    xr = None if proj_xr is None else np.ascontiguousarray(proj_xr, np.float32)
    ur = None if u_right2 is None else np.ascontiguousarray(u_right2, np.float32)
    mc = np.full(max(len(k2), 1), -1, np.int32)
    f = getattr(L, fn); f.restype = C.c_int
    n = f(_p(p), _p(xr), _p(dm), C.c_int(len(p)), _p(k2), _p(d2), _p(hd), _p(ur), C.c_int(len(k2)), _p(b), _p(sf), C.c_int(len(sf)),
          C.c_float(th), C.c_int(int(far_points)), C.c_float(th_far), C.c_float(nnratio), _p(mc))
    return n, mc[:len(k2)].copy()


def search_by_projection_map_points(pts, descMP, kps2, desc2, held2, bounds, scale_factors, th=1.0, far_points=False, th_far=0.0,
                                    nnratio=0.8):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints), monocular -> (nmatches, match_cur[n2])"""
    from eorb_slam_b200.synth import TRACK_POINT_DTYPE
    L = lib()
    p = np.ascontiguousarray(pts, TRACK_POINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
    dm = np.ascontiguousarray(descMP, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    hd = None if held2 is None else np.ascontiguousarray(held2, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    mc = np.full(max(len(k2), 1), -1, np.int32)
    L.orc_search_by_projection_map_points.restype = C.c_int
    n = L.orc_search_by_projection_map_points(_p(p), _p(dm), C.c_int(len(p)), _p(k2), _p(d2), _p(hd) if hd is not None else None,
                                              C.c_int(len(k2)), _p(b), _p(sf), C.c_int(len(sf)), C.c_float(th), C.c_int(int(far_points)),
                                              C.c_float(th_far), C.c_float(nnratio), _p(mc))
    return n, mc[:len(k2)].copy()


def search_by_bow(kps_kf, desc_kf, valid_kf, fv_kf, kps_f, desc_f, fv_f, nnratio=0.7, check_ori=True):
    """ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches), monocular; fv_* = (nodes, start, feats) CSR -> (nmatches, match_f[n2])"""
    L = lib()
    k1 = np.ascontiguousarray(kps_kf, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps_f, KEYPOINT_DTYPE)
    d1 = np.ascontiguousarray(desc_kf, np.uint8); d2 = np.ascontiguousarray(desc_f, np.uint8); v = np.ascontiguousarray(valid_kf, np.uint8)
    a = [np.ascontiguousarray(fv_kf[0], np.uint32), np.ascontiguousarray(fv_kf[1], np.int32), np.ascontiguousarray(fv_kf[2], np.uint32)]
    b = [np.ascontiguousarray(fv_f[0], np.uint32), np.ascontiguousarray(fv_f[1], np.int32), np.ascontiguousarray(fv_f[2], np.uint32)]
    mf = np.full(max(len(k2), 1), -1, np.int32)
    L.orc_search_by_bow.restype = C.c_int
    n = L.orc_search_by_bow(_p(k1), _p(d1), _p(v), _p(a[0]), _p(a[1]), _p(a[2]), C.c_int(len(a[0])), _p(k2), _p(d2), C.c_int(len(k2)),
                            _p(b[0]), _p(b[1]), _p(b[2]), C.c_int(len(b[0])), C.c_float(nnratio), C.c_int(int(check_ori)), _p(mf))
    return n, mf[:len(k2)].copy()


def search_by_bow_kf(kps1, desc1, valid1, fv1, kps2, desc2, valid2, fv2, nnratio=0.7, check_ori=True, _fn=None):
    """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:833-990), monocular keyframes -> (nmatches, match12[n1])"""
    k1 = np.ascontiguousarray(kps1, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
    d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    v1 = np.ascontiguousarray(valid1, np.uint8); v2 = np.ascontiguousarray(valid2, np.uint8)
    a = [np.ascontiguousarray(fv1[0], np.uint32), np.ascontiguousarray(fv1[1], np.int32), np.ascontiguousarray(fv1[2], np.uint32)]
    b = [np.ascontiguousarray(fv2[0], np.uint32), np.ascontiguousarray(fv2[1], np.int32), np.ascontiguousarray(fv2[2], np.uint32)]
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    f = _fn or lib().orc_search_by_bow_kf
    f.restype = C.c_int
    n = f(_p(k1), _p(d1), _p(v1), _p(a[0]), _p(a[1]), _p(a[2]), C.c_int(len(a[0])), C.c_int(len(k1)), _p(k2), _p(d2), _p(v2), _p(b[0]), _p(b[1]),
          _p(b[2]), C.c_int(len(b[0])), C.c_int(len(k2)), C.c_float(nnratio), C.c_int(int(check_ori)), _p(m12))
    return n, m12[:len(k1)].copy()


# ---- bag of words + undistortion
class VocabOracle:
    def __init__(self, voc):
        """voc: dict from synth.make_vocabulary (k, L, scoring, weighting, parent, is_leaf, desc, weight)"""
        L = lib()
        L.orc_vocab_create.restype = C.c_void_p
        self.voc = voc
        self.h = C.c_void_p(L.orc_vocab_create(C.c_int(voc["k"]), C.c_int(voc["L"]), C.c_int(voc["scoring"]), C.c_int(voc["weighting"]),
                                               C.c_int(len(voc["parent"])), _p(voc["parent"]), _p(voc["is_leaf"]), _p(voc["desc"]),
                                               _p(voc["weight"])))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_vocab_destroy(self.h)
            self.h = None

    def transform(self, feats, levelsup=4):
        """-> dict(word_id, word_w, node_id, bow_ids, bow_vals, fv_nodes, fv_start, fv_feats)"""
        f = np.ascontiguousarray(feats, np.uint8).reshape(-1, 32)
        n = len(f); m = max(n, 1)
        wid = np.zeros(m, np.uint32); ww = np.zeros(m, np.float64); nid = np.zeros(m, np.uint32)
        bi = np.zeros(m, np.uint32); bv = np.zeros(m, np.float64); fn = np.zeros(m, np.uint32); fs = np.zeros(m + 1, np.int32)
        ff = np.zeros(m, np.uint32); nb = C.c_int(0); nf = C.c_int(0)
        lib().orc_vocab_transform(self.h, _p(f), C.c_int(n), C.c_int(levelsup), _p(wid), _p(ww), _p(nid), _p(bi), _p(bv), C.byref(nb),
                                  _p(fn), _p(fs), _p(ff), C.byref(nf))
        nbow, nfv = nb.value, nf.value
        return dict(word_id=wid[:n], word_w=ww[:n], node_id=nid[:n], bow_ids=bi[:nbow].copy(), bow_vals=bv[:nbow].copy(),
                    fv_nodes=fn[:nfv].copy(), fv_start=fs[:nfv + 1].copy(), fv_feats=ff[:fs[nfv]].copy())


def undistort_points(xy, K4, dist5):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.zeros_like(xy)
    K = np.ascontiguousarray(K4, np.float32); D = np.ascontiguousarray(dist5, np.float32)
    lib().orc_undistort_points(_p(xy), C.c_int(len(xy)), _p(K), _p(D), _p(out))
    return out
