"""ctypes host mirror of the reference's class surface over the C ABI (include/eorb_b200.h).

The reference is C++; the drop-in C++ shims live in eorb_slam_b200/shim/.  This module gives tests and bench.py
the same names and argument meaning from Python:

    ORBextractor(ORBxParams)(image, mask, vLappingArea)          include/ORBextractor.h:62-136
    ORBmatcher(nnratio, checkOri).DescriptorDistance / best-2     include/ORBmatcher.h:35-114
    EvImConverter.ev2im / ev2im_gauss / ev2mci_gg_f               include/Event/EventConversion.h:40-81

There is no CPU fallback: importing this module raises if libeorb_b200.so is missing, and every compute call
raises EorbError when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .synth import EVENT_DTYPE, KEYPOINT_DTYPE, MATCH_DTYPE, TRACK_POINT_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
# EORB_B200_LIB selects another build of the same library (kernel-tuning A/B runs); there is still no CPU fallback
LIB_PATH = os.environ.get("EORB_B200_LIB") or os.path.join(_HERE, "libeorb_b200.so")
BEST2_DTYPE = np.dtype([("key1", "<u8"), ("key2", "<u8")])

EORB_OK, EORB_EMPTY = 0, -1
EV_NEAREST, EV_GAUSS, EV_SE3, EV_SE2 = 0, 1, 2, 3
NORM_NONE, NORM_RUNNING, NORM_MINMAX = 0, 1, 2


class EorbError(RuntimeError):
    pass


class _OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scaleFactor", C.c_float), ("nlevels", C.c_int), ("iniThFAST", C.c_int),
                ("minThFAST", C.c_int), ("edgeTh", C.c_int), ("imW", C.c_int), ("imH", C.c_int)]


class _EvParams(C.Structure):
    _fields_ = [("mode", C.c_int), ("width", C.c_int), ("height", C.c_int), ("sigma", C.c_float), ("pol", C.c_int),
                ("normalize", C.c_int), ("Tcw", C.c_float * 16), ("med_depth", C.c_float), ("K", C.c_float * 4),
                ("se2", C.c_float * 4), ("se2_n", C.c_int), ("cam_model", C.c_int), ("kb", C.c_float * 4)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libeorb_b200.so not built (run `python -c 'import __graft_entry__ as g; g.build()'`): "
                          "eorb_slam_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i, f, sz, i64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64
    sig = {
        "eorb_version": ([], i), "eorb_last_error": ([], C.c_char_p), "eorb_device_count": ([], i),
        "eorb_timer_create": ([C.POINTER(vp)], i), "eorb_timer_destroy": ([vp], i), "eorb_timer_start": ([vp, vp], i),
        "eorb_timer_stop": ([vp, vp], i), "eorb_timer_elapsed_ms": ([vp, C.POINTER(f)], i),
        "eorb_probe_popc_rate": ([i, C.POINTER(C.c_double)], i), "eorb_selftest_math": ([i, C.POINTER(i)], i),
        "eorb_orb_create": ([C.POINTER(_OrbParams), i, i, C.POINTER(vp)], i), "eorb_orb_destroy": ([vp], i),
        "eorb_orb_set_stream": ([vp, vp], i), "eorb_orb_reset_stream": ([vp], i), "eorb_orb_get_stream": ([vp], vp), "eorb_orb_synchronize": ([vp], i),
        "eorb_orb_tables": ([vp, vp, vp, vp, vp, vp, vp, vp], i), "eorb_orb_params_tables": ([vp, vp, vp, vp, vp, vp, vp, vp], i), "eorb_orb_max_keypoints": ([vp], i), "eorb_orb_max_keypoints_for_size": ([vp, i, i], i),
        "eorb_orb_extract": ([vp, vp, i, i, sz, i, i, i, vp, vp, i, vp], i),
        "eorb_orb_extract_batch": ([vp, vp, i, i, i, sz, sz, i, i, i, vp, vp, i, vp, vp], i),
        "eorb_orb_extract_batch_device": ([vp, vp, i, i, i, sz, sz, i, i, i, vp, vp, i, vp, vp], i),
        "eorb_orb_level_size": ([vp, i, vp, vp], i), "eorb_orb_pyramid_level": ([vp, i, i, vp, sz], i),
        "eorb_orb_tracked_desc": ([vp, vp, i, i, sz, vp, i, vp], i),
        "eorb_orb_assign_level_by_best_desc": ([vp, vp, vp, i, i, sz, vp, i], i),
        "eorb_orb_debug_blurred": ([vp, i, i, vp, sz], i), "eorb_orb_debug_candidates": ([vp, i, i, vp, vp, vp, i], i),
        "eorb_orb_debug_level_kps": ([vp, i, i, vp, vp, vp, vp, i], i), "eorb_orb_launch_count": ([vp], C.c_longlong),
        "eorb_orb_stage_timing": ([vp, i], i), "eorb_orb_stage_times": ([vp, vp, vp], i),
        "eorb_descriptor_distance": ([vp, vp], i),
        "eorb_matcher_create": ([i, C.POINTER(vp)], i), "eorb_matcher_destroy": ([vp], i),
        "eorb_matcher_set_stream": ([vp, vp], i), "eorb_matcher_reset_stream": ([vp], i), "eorb_matcher_synchronize": ([vp], i),
        "eorb_matcher_launch_count": ([vp], C.c_longlong),
        "eorb_matcher_set_db_host": ([vp, vp, i64, i64], i), "eorb_matcher_set_db_device": ([vp, vp, i64, i64], i),
        "eorb_matcher_set_engine": ([vp, i], i), "eorb_matcher_last_engine": ([vp], i),
        "eorb_matcher_search": ([vp, vp, i, i, f, vp], i), "eorb_matcher_search_device": ([vp, vp, i, vp], i),
        "eorb_matcher_merge_device": ([vp, vp, i, i, i, f, vp], i),
        "eorb_matcher_search_sharded": ([vp, vp, i, i, f, vp, i, vp], i),
        "eorb_nccl_unique_id": ([vp], i), "eorb_nccl_comm_init_rank": ([C.POINTER(vp), i, vp, i, i], i), "eorb_nccl_comm_destroy": ([vp], i),
        "eorb_hamming_best2": ([vp, i, vp, i64, i, f, vp], i), "eorb_rotation_filter": ([vp, vp, vp, i], i),
        "eorb_ev_create": ([i, i, i64, i, i, C.POINTER(vp)], i), "eorb_ev_destroy": ([vp], i),
        "eorb_ev_set_stream": ([vp, vp], i), "eorb_ev_reset_stream": ([vp], i), "eorb_ev_synchronize": ([vp], i), "eorb_ev_launch_count": ([vp], C.c_longlong),
        "eorb_ev_accumulate": ([vp, vp, i64, C.POINTER(_EvParams), vp, vp, vp], i),
        "eorb_ev_accumulate_batch_device": ([vp, vp, vp, i, C.POINTER(_EvParams), vp, vp, vp], i),
        "eorb_ev_accumulate_batch": ([vp, vp, vp, i, C.POINTER(_EvParams), vp, vp, vp], i),
        "eorb_ev_mci_jac": ([vp, vp, i64, i, i, f, vp, f, vp, i, i, vp], i),
        "eorb_ev_image_focus_device": ([vp, vp, i, i, i, i, i, vp], i), "eorb_ev_image_focus": ([vp, vp, i, i, sz, i, i, vp], i),
        "eorb_lk_create": ([i, i, i, i, C.POINTER(vp)], i), "eorb_lk_destroy": ([vp], i),
        "eorb_lk_set_stream": ([vp, vp], i), "eorb_lk_reset_stream": ([vp], i), "eorb_lk_launch_count": ([vp], C.c_longlong),
        "eorb_lk_set_ref": ([vp, vp, i, i, sz, vp, i, i, i], i), "eorb_lk_set_ref_device": ([vp, vp, i, i, sz, vp, i, i, i], i),
        "eorb_lk_track": ([vp, vp, sz, vp, i, C.c_double, f, vp, vp, vp], i),
        "eorb_lk_track_device": ([vp, vp, sz, vp, i, C.c_double, f, vp, vp, vp], i),
        "eorb_lk_set_ref_keypoints": ([vp, vp, i, i, sz, i, vp, i, i, i], i), "eorb_lk_set_last_tracked": ([vp, vp, i], i),
        "eorb_lk_get_last_tracked": ([vp, vp], i),
        "eorb_lk_track_and_match": ([vp, vp, sz, i, i, C.c_double, f, i, vp, vp, vp, vp], i),
        "eorb_lk_track_and_match_device": ([vp, vp, sz, i, C.c_double, f, i, vp, vp, vp, vp], i),
        "eorb_guided_create": ([i, C.POINTER(vp)], i), "eorb_guided_destroy": ([vp], i),
        "eorb_guided_set_stream": ([vp, vp], i), "eorb_guided_reset_stream": ([vp], i), "eorb_guided_launch_count": ([vp], C.c_longlong),
        "eorb_guided_frame_grid": ([vp, vp, i, vp, vp, vp, vp], i),
        "eorb_guided_features_in_area": ([vp, vp, i, vp, vp, i, vp, vp, i], i),
        "eorb_guided_search_for_initialization": ([vp, vp, vp, i, vp, vp, i, vp, vp, i, f, i, vp, vp], i),
        "eorb_guided_search_for_initialization_device": ([vp, vp, vp, i, vp, vp, i, vp, vp, i, f, i, vp, vp], i),
        "eorb_guided_search_by_projection": ([vp, vp, vp, vp, vp, vp, i, vp, vp, i, vp, vp, vp, i, f, i, vp, vp], i),
        "eorb_guided_search_by_projection_device": ([vp, vp, vp, vp, vp, vp, i, vp, vp, i, vp, vp, vp, i, f, i, vp, vp], i),
        "eorb_guided_search_by_projection_stereo": ([vp, vp, vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, f, i, i, f, vp, vp], i),
        "eorb_guided_search_by_projection_stereo_device": ([vp, vp, vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, f, i, i, f, vp, vp], i),
        "eorb_guided_search_by_projection_reloc": ([vp, vp, vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, f, i, i, vp, vp], i),
        "eorb_guided_search_by_projection_reloc_device": ([vp, vp, vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, f, i, i, vp, vp], i),
        "eorb_guided_search_by_projection_map_points_stereo": ([vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, i, f, i, f, f, vp, vp], i),
        "eorb_guided_search_by_projection_map_points_stereo_device": ([vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, i, f, i, f, f, vp, vp], i),
        "eorb_guided_search_by_bow": ([vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, i, vp, vp, vp, i, f, i, vp, vp], i),
        "eorb_guided_search_by_bow_device": ([vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, i, vp, vp, vp, i, f, i, vp, vp], i),
        "eorb_guided_search_for_triangulation": ([vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, vp, i, i, i, vp, vp], i),
        "eorb_guided_search_for_triangulation_device": ([vp, vp, vp, vp, i, vp, vp, vp, i, i, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, vp, i, i, i, vp, vp], i),
        "eorb_guided_search_windows": ([vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, vp, i, i, i, vp, vp, vp, vp], i),
        "eorb_guided_search_windows_device": ([vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, vp, i, i, i, vp, vp, vp, vp], i),
        "eorb_guided_search_by_bow_kf": ([vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, f, i, vp, vp], i),
        "eorb_guided_search_by_bow_kf_device": ([vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, f, i, vp, vp, vp], i),
        "eorb_guided_search_by_projection_map_points": ([vp, vp, vp, i, vp, vp, vp, i, vp, vp, i, f, i, f, f, vp, vp], i),
        "eorb_guided_search_by_projection_map_points_device": ([vp, vp, vp, i, vp, vp, vp, i, vp, vp, i, f, i, f, f, vp, vp], i),
        "eorb_vocab_create": ([i, i, i, i, i, i, vp, vp, vp, vp, C.POINTER(vp)], i), "eorb_vocab_destroy": ([vp], i),
        "eorb_vocab_set_stream": ([vp, vp], i), "eorb_vocab_reset_stream": ([vp], i), "eorb_vocab_launch_count": ([vp], C.c_longlong),
        "eorb_vocab_transform": ([vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp], i),
        "eorb_vocab_transform_device": ([vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp], i),
        "eorb_vocab_transform_resident": ([vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp], i),
        "eorb_undistort_keypoints": ([vp, i, vp, vp, vp], i), "eorb_undistort_keypoints_device": ([vp, vp, i, vp, vp, vp], i),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)   # AttributeError here == header/library mismatch
        fn.argtypes = args
        fn.restype = res
    return L, list(sig)


lib, EXPORTED = _load()


def nccl_unique_id() -> np.ndarray:
    """128-byte ncclUniqueId (rank 0 creates it, every rank passes the same bytes to nccl_comm_init_rank)"""
    out = np.zeros(128, np.uint8)
    _check(lib.eorb_nccl_unique_id(out.ctypes.data_as(C.c_void_p)), "nccl_unique_id")
    return out


def nccl_comm_init_rank(nranks: int, unique_id: np.ndarray, rank: int, device: int) -> int:
    comm = C.c_void_p()
    uid = np.ascontiguousarray(unique_id, np.uint8)
    _check(lib.eorb_nccl_comm_init_rank(C.byref(comm), nranks, uid.ctypes.data_as(C.c_void_p), rank, device), "nccl_comm_init_rank")
    return comm.value


def nccl_comm_destroy(comm: int):
    _check(lib.eorb_nccl_comm_destroy(C.c_void_p(int(comm))), "nccl_comm_destroy")


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


def _check(rc: int, what: str) -> int:
    if rc <= -2:
        raise EorbError("%s: error %d: %s" % (what, rc, lib.eorb_last_error().decode(errors="replace")))
    return rc


def device_count() -> int:
    return lib.eorb_device_count()


def probe_popc_rate(device: int = 0) -> float:
    r = C.c_double(0)
    _check(lib.eorb_probe_popc_rate(device, C.byref(r)), "probe_popc_rate")
    return r.value


def selftest_math(device: int = 0) -> int:
    n = C.c_int(-1)
    _check(lib.eorb_selftest_math(device, C.byref(n)), "selftest_math")
    return n.value


class CudaTimer:
    """cudaEvent pair recorded on a given stream (the stream the kernels are launched on)."""

    def __init__(self):
        self.h = C.c_void_p()
        _check(lib.eorb_timer_create(C.byref(self.h)), "timer_create")

    def start(self, stream):
        _check(lib.eorb_timer_start(self.h, C.c_void_p(int(stream or 0))), "timer_start")

    def stop(self, stream):
        _check(lib.eorb_timer_stop(self.h, C.c_void_p(int(stream or 0))), "timer_stop")

    def elapsed_ms(self) -> float:
        ms = C.c_float(0)
        _check(lib.eorb_timer_elapsed_ms(self.h, C.byref(ms)), "timer_elapsed")
        return ms.value

    def __del__(self):
        try:
            lib.eorb_timer_destroy(self.h)
        except Exception:
            pass


# ================================================================================================ ORB
@dataclass
class ORBxParams:
    """ORB_SLAM3::ORBxParams (include/ORBextractor.h:33-47)"""
    nfeatures: int = 1000
    scaleFactor: float = 1.2
    nlevels: int = 8
    iniThFAST: int = 20
    minThFAST: int = 7
    edgeTh: int = 19
    imSize: tuple = (752, 480)


class ORBextractor:
    def __init__(self, params: ORBxParams, device: int = 0, max_batch: int = 1):
        self.params = params
        cp = _OrbParams(params.nfeatures, params.scaleFactor, params.nlevels, params.iniThFAST, params.minThFAST,
                        params.edgeTh, params.imSize[0], params.imSize[1])
        self.h = C.c_void_p()
        _check(lib.eorb_orb_create(C.byref(cp), device, max_batch, C.byref(self.h)), "ORBextractor")
        self.max_batch = max_batch
        self.nlevels = params.nlevels
        self.cap = lib.eorb_orb_max_keypoints(self.h)
        self._last = None

    def __del__(self):
        try:
            if self.h:
                lib.eorb_orb_destroy(self.h); self.h = None
        except Exception:
            pass

    # --- reference getters (ORBextractor.h:83-107)
    def _tables(self):
        nl = self.nlevels
        n = C.c_int(); e = C.c_int()
        arrs = [np.empty(nl, np.float32) for _ in range(4)]
        fpl = np.empty(nl, np.int32)
        _check(lib.eorb_orb_tables(self.h, C.byref(n), C.byref(e), *[_p(a) for a in arrs], _p(fpl)), "tables")
        return n.value, e.value, arrs, fpl

    def GetLevels(self): return self._tables()[0]
    def GetScaleFactor(self): return float(np.float32(self.params.scaleFactor))
    def GetScaleFactors(self): return self._tables()[2][0]
    def GetInverseScaleFactors(self): return self._tables()[2][1]
    def GetScaleSigmaSquares(self): return self._tables()[2][2]
    def GetInverseScaleSigmaSquares(self): return self._tables()[2][3]
    def GetNumFeatures(self): return self.params.nfeatures
    def edge_threshold(self): return self._tables()[1]
    def features_per_level(self): return self._tables()[3]

    def set_stream(self, stream):
        """adopt a cudaStream_t given as an int (0 = CUDA default stream); None returns to the handle's own stream"""
        if stream is None:
            _check(lib.eorb_orb_reset_stream(self.h), "reset_stream")
        else:
            _check(lib.eorb_orb_set_stream(self.h, C.c_void_p(int(stream))), "set_stream")
    def stream(self): return lib.eorb_orb_get_stream(self.h)
    def synchronize(self): _check(lib.eorb_orb_synchronize(self.h), "synchronize")
    def launch_count(self): return lib.eorb_orb_launch_count(self.h)

    STAGES = ("pyramid", "fast", "octree", "index", "blur", "orient_desc")

    def stage_timing(self, enable=True): _check(lib.eorb_orb_stage_timing(self.h, int(enable)), "stage_timing")

    def stage_times(self):
        ms = np.zeros(6, np.float32); ln = np.zeros(6, np.int64)
        _check(lib.eorb_orb_stage_times(self.h, _p(ms), _p(ln)), "stage_times")
        return dict(zip(self.STAGES, ms.tolist())), dict(zip(self.STAGES, ln.tolist()))

    # --- operator() (ORBextractor.cc:1092-1238)
    def __call__(self, image, mask=None, vLappingArea=(0, 1000), want_desc=True):
        """-> (ret, keypoints[KEYPOINT_DTYPE], descriptors (n,32) u8 | None).  ret = monoIndex, -1 for an empty image."""
        if image is None or np.size(image) == 0:
            return EORB_EMPTY, np.empty(0, KEYPOINT_DTYPE), None
        img = np.ascontiguousarray(image, np.uint8)
        assert img.ndim == 2, "CV_8UC1 expected (ORBextractor.cc:1100)"
        cap = lib.eorb_orb_max_keypoints_for_size(self.h, img.shape[1], img.shape[0])   # elongated frames have more octree roots
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8) if want_desc else None
        n = C.c_int(0)
        ret = _check(lib.eorb_orb_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], int(vLappingArea[0]),
                                          int(vLappingArea[1]), int(want_desc), _p(kps), _p(desc), cap, C.byref(n)), "extract")
        return ret, kps[:n.value].copy(), (desc[:n.value].copy() if want_desc else None)

    def extract_batch(self, frames, vLappingArea=(0, 1000), want_desc=True, out=None):
        """frames: (n,h,w) u8 host array.  -> (kps (n,cap), desc (n,cap,32), n_out (n,), mono (n,))"""
        frames = np.ascontiguousarray(frames, np.uint8) if isinstance(frames, np.ndarray) else frames
        n, h, w = frames.shape
        if out is None:
            kps = np.zeros((n, self.cap), KEYPOINT_DTYPE)
            desc = np.zeros((n, self.cap, 32), np.uint8) if want_desc else None
            nout = np.zeros(n, np.int32); mono = np.zeros(n, np.int32)
        else:
            kps, desc, nout, mono = out
        _check(lib.eorb_orb_extract_batch(self.h, _p(frames), n, w, h, w, w * h, int(vLappingArea[0]), int(vLappingArea[1]),
                                          int(want_desc), _p(kps), _p(desc), self.cap, _p(nout), _p(mono)), "extract_batch")
        return kps, desc, nout, mono

    def extract_batch_raw(self, imgs_ptr, n, w, h, row_stride, frame_stride, lap, want_desc, kps_ptr, desc_ptr, cap, n_ptr, mono_ptr,
                          device=False):
        """raw-pointer form (host or device pointers); used by bench.py with pinned / resident buffers"""
        fn = lib.eorb_orb_extract_batch_device if device else lib.eorb_orb_extract_batch
        return _check(fn(self.h, _p(imgs_ptr), n, w, h, row_stride, frame_stride, int(lap[0]), int(lap[1]), int(want_desc),
                         _p(kps_ptr), _p(desc_ptr), cap, _p(n_ptr), _p(mono_ptr)), "extract_batch_raw")

    # --- mvImagePyramid (ORBextractor.h:105)
    def level_size(self, level):
        w = C.c_int(); h = C.c_int()
        _check(lib.eorb_orb_level_size(self.h, level, C.byref(w), C.byref(h)), "level_size")
        return w.value, h.value

    def pyramid_level(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        _check(lib.eorb_orb_pyramid_level(self.h, frame, level, _p(out), out.strides[0]), "pyramid_level")
        return out

    @property
    def mvImagePyramid(self):
        return [self.pyramid_level(l) for l in range(self.nlevels)]

    # --- stage taps (tests)
    def debug_blurred(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        _check(lib.eorb_orb_debug_blurred(self.h, frame, level, _p(out), out.strides[0]), "debug_blurred")
        return out

    def debug_candidates(self, level, frame=0, cap=1 << 20):
        xs = np.empty(cap, np.int32); ys = np.empty(cap, np.int32); sc = np.empty(cap, np.int32)
        n = _check(lib.eorb_orb_debug_candidates(self.h, frame, level, _p(xs), _p(ys), _p(sc), cap), "debug_candidates")
        return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()

    def debug_level_kps(self, level, frame=0):
        cap = self.cap + 64
        xs = np.empty(cap, np.int32); ys = np.empty(cap, np.int32); sc = np.empty(cap, np.int32); an = np.empty(cap, np.float32)
        n = _check(lib.eorb_orb_debug_level_kps(self.h, frame, level, _p(xs), _p(ys), _p(sc), _p(an), cap), "debug_level_kps")
        return xs[:n].copy(), ys[:n].copy(), sc[:n].copy(), an[:n].copy()

    # --- secondary API (ORBextractor.cc:1267-1363)
    def ComputeTrackedKPtsDesc(self, trackedImage, trackedKPts):
        img = np.ascontiguousarray(trackedImage, np.uint8)
        kps = np.ascontiguousarray(trackedKPts, KEYPOINT_DTYPE)
        desc = np.zeros((len(kps), 32), np.uint8)
        _check(lib.eorb_orb_tracked_desc(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), len(kps), _p(desc)),
               "ComputeTrackedKPtsDesc")
        return desc

    def AssignKPtLevelByBestDesc(self, refDescs, trackedImage, trackedKPts):
        img = np.ascontiguousarray(trackedImage, np.uint8)
        kps = np.ascontiguousarray(trackedKPts, KEYPOINT_DTYPE).copy()
        ref = np.ascontiguousarray(refDescs, np.uint8)
        _check(lib.eorb_orb_assign_level_by_best_desc(self.h, _p(ref), _p(img), img.shape[1], img.shape[0], img.strides[0],
                                                      _p(kps), len(kps)), "AssignKPtLevelByBestDesc")
        return kps


# ================================================================================================ matcher
class ORBmatcher:
    TH_LOW = 50      # ORBmatcher.cc:36-38
    TH_HIGH = 100
    HISTO_LENGTH = 30

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, device: int = 0):
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self.h = C.c_void_p()
        _check(lib.eorb_matcher_create(device, C.byref(self.h)), "ORBmatcher")
        self._db_keep = None

    def __del__(self):
        try:
            if self.h:
                lib.eorb_matcher_destroy(self.h); self.h = None
        except Exception:
            pass

    @staticmethod
    def DescriptorDistance(a, b) -> int:
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return _check(lib.eorb_descriptor_distance(_p(a), _p(b)), "DescriptorDistance")

    def set_stream(self, stream):
        if stream is None:
            _check(lib.eorb_matcher_reset_stream(self.h), "reset_stream")
        else:
            _check(lib.eorb_matcher_set_stream(self.h, C.c_void_p(int(stream))), "set_stream")
    def synchronize(self): _check(lib.eorb_matcher_synchronize(self.h), "synchronize")
    def launch_count(self): return lib.eorb_matcher_launch_count(self.h)

    HAMMING_POPC, HAMMING_TENSOR, HAMMING_AUTO = 0, 1, 2

    def set_engine(self, engine):
        """scan engine: POPC (integer pipe), TENSOR (tcgen05 int8 contraction, bit-identical results) or AUTO"""
        _check(lib.eorb_matcher_set_engine(self.h, int(engine)), "set_engine")

    def last_engine(self): return lib.eorb_matcher_last_engine(self.h)

    def set_db(self, db, index_offset=0):
        db = np.ascontiguousarray(db, np.uint8)
        assert db.ndim == 2 and db.shape[1] == 32
        _check(lib.eorb_matcher_set_db_host(self.h, _p(db), len(db), index_offset), "set_db")

    def set_db_device(self, d_ptr, ndb, index_offset=0):
        _check(lib.eorb_matcher_set_db_device(self.h, _p(d_ptr), ndb, index_offset), "set_db_device")

    def search(self, q, th=None):
        """best-2 + TH + ratio over the resident database -> MATCH_DTYPE[nq]"""
        q = np.ascontiguousarray(q, np.uint8)
        out = np.zeros(len(q), MATCH_DTYPE)
        _check(lib.eorb_matcher_search(self.h, _p(q), len(q), self.TH_LOW if th is None else th, self.mfNNratio, _p(out)), "search")
        return out

    def search_raw(self, q_ptr, nq, out_ptr, th=None):
        return _check(lib.eorb_matcher_search(self.h, _p(q_ptr), nq, self.TH_LOW if th is None else th, self.mfNNratio, _p(out_ptr)), "search")

    def search_device(self, d_q, nq, d_partial):
        _check(lib.eorb_matcher_search_device(self.h, _p(d_q), nq, _p(d_partial)), "search_device")

    def merge_device(self, d_gathered, nshards, nq, d_out, th=None):
        _check(lib.eorb_matcher_merge_device(self.h, _p(d_gathered), nshards, nq, self.TH_LOW if th is None else th, self.mfNNratio,
                                             _p(d_out)), "merge_device")

    def search_sharded(self, d_q, nq, comm, nshards, d_out, th=None):
        """this rank's shard scan + ncclAllGather + merge, all on the matcher's stream (asynchronous)"""
        _check(lib.eorb_matcher_search_sharded(self.h, _p(d_q), nq, self.TH_LOW if th is None else th, self.mfNNratio,
                                               C.c_void_p(int(comm)), nshards, _p(d_out)), "search_sharded")

    def SearchBruteForce(self, desc1, desc2, angles1=None, angles2=None, th=None):
        """Frame-to-frame brute force in the shape of Frame.cc:1228-1235 with the ORBmatcher acceptance rule and,
        if checkOri and angles are given, the rotation-histogram filter (ORBmatcher.cc:784-823).
        -> (nmatches, vnMatches12 int32[n1])"""
        self.set_db(desc2)
        m = self.search(desc1, th)
        match12 = np.where(m["accepted"] == 1, m["best_idx"], -1).astype(np.int32)
        n = int((match12 >= 0).sum())
        if self.mbCheckOrientation and angles1 is not None and angles2 is not None:
            a1 = np.ascontiguousarray(angles1, np.float32); a2 = np.ascontiguousarray(angles2, np.float32)
            n = _check(lib.eorb_rotation_filter(_p(a1), _p(a2), _p(match12), len(match12)), "rotation_filter")
        return n, match12


def rotation_filter(angle1, angle2, match12):
    a1 = np.ascontiguousarray(angle1, np.float32); a2 = np.ascontiguousarray(angle2, np.float32)
    m = np.ascontiguousarray(match12, np.int32).copy()
    n = _check(lib.eorb_rotation_filter(_p(a1), _p(a2), _p(m), len(m)), "rotation_filter")
    return n, m


def hamming_best2(q, db, th=50, ratio=0.7):
    q = np.ascontiguousarray(q, np.uint8); db = np.ascontiguousarray(db, np.uint8)
    out = np.zeros(len(q), MATCH_DTYPE)
    _check(lib.eorb_hamming_best2(_p(q), len(q), _p(db), len(db), th, ratio, _p(out)), "hamming_best2")
    return out


# ================================================================================================ events
class EvImConverter:
    """EORB_SLAM::EvImConverter (static methods in the reference; here bound to a converter handle that owns the
    device workspace).  `camera` = (fx, fy, cx, cy) of the Pinhole model."""

    def __init__(self, device=0, max_windows=1, max_events=1 << 20, max_width=346, max_height=260):
        self.h = C.c_void_p()
        _check(lib.eorb_ev_create(device, max_windows, max_events, max_width, max_height, C.byref(self.h)), "EvImConverter")

    def __del__(self):
        try:
            if self.h:
                lib.eorb_ev_destroy(self.h); self.h = None
        except Exception:
            pass

    def set_stream(self, stream):
        if stream is None:
            _check(lib.eorb_ev_reset_stream(self.h), "reset_stream")
        else:
            _check(lib.eorb_ev_set_stream(self.h, C.c_void_p(int(stream))), "set_stream")
    def synchronize(self): _check(lib.eorb_ev_synchronize(self.h), "synchronize")
    def launch_count(self): return lib.eorb_ev_launch_count(self.h)

    @staticmethod
    def make_params(mode, w, h, sigma=1.0, pol=False, normalize=NORM_NONE, Tcw=None, medDepth=1.0, camera=None, params2D=None):
        p = _EvParams()
        p.mode = mode; p.width = w; p.height = h; p.sigma = sigma; p.pol = int(pol); p.normalize = normalize
        T = np.eye(4, dtype=np.float32) if Tcw is None else np.asarray(Tcw, np.float32).reshape(4, 4)
        for k, v in enumerate(T.reshape(16)):
            p.Tcw[k] = float(v)
        p.med_depth = medDepth
        cam = (1.0, 1.0, 0.0, 0.0) if camera is None else camera
        for k in range(4):
            p.K[k] = float(cam[k])
        p.cam_model = 0
        if len(cam) >= 8:                      # a KannalaBrandt8 camera carries fx, fy, cx, cy, k1..k4 (KannalaBrandt8.h mvParameters)
            p.cam_model = 1
            for k in range(4):
                p.kb[k] = float(cam[4 + k])
        p.se2_n = 0
        if params2D is not None:
            s = np.asarray(params2D, np.float32).reshape(-1)
            p.se2_n = len(s)
            for k in range(len(s)):
                p.se2[k] = float(s[k])
        return p

    def _run(self, evs, p):
        evs = np.ascontiguousarray(evs)
        assert evs.dtype.itemsize == 24, "EventData is a 24-byte record (include/Event/EventData.h:36-58)"
        img = np.zeros((p.height, p.width), np.float32)
        u8 = np.zeros((p.height, p.width), np.uint8) if p.normalize != NORM_NONE else None
        mm = np.zeros(2, np.float32)
        rc = _check(lib.eorb_ev_accumulate(self.h, _p(evs) if len(evs) else None, len(evs), C.byref(p), _p(img), _p(u8), _p(mm)), "ev_accumulate")
        return rc, img, u8, mm

    def ev2im(self, vEvData, imWidth, imHeight, pol=False, normalized=True):
        p = self.make_params(EV_NEAREST, imWidth, imHeight, 1.0, pol, NORM_RUNNING if normalized else NORM_NONE)
        _, img, u8, _ = self._run(vEvData, p)
        return u8 if normalized else img

    def ev2mci_gg_f_jac(self, vEvData, camera, R, t, medDepth, imWidth, imHeight, sigma, pol=False, global_mean=False):
        """EvImConverter::ev2mci_gg_f_jac (EventConversion.cc:533-662): R (3x3), t (3) = the SE3 vertex estimate -> float64[6]"""
        ev = np.ascontiguousarray(vEvData)
        Rt = np.concatenate([np.asarray(R, np.float64).reshape(9), np.asarray(t, np.float64).reshape(3)])
        K = np.ascontiguousarray(camera, np.float32)
        out = np.zeros(6, np.float64)
        _check(lib.eorb_ev_mci_jac(self.h, _p(ev), len(ev), imWidth, imHeight, float(sigma), _p(Rt), float(medDepth), _p(K), int(pol),
                                   int(global_mean), _p(out)), "ev2mci_gg_f_jac")
        return out

    # ---- contrast metric (EventConversion.cc:74-162)
    FOCUS_LOCAL_STD, FOCUS_GLOBAL_STD, FOCUS_LOCAL_MEAN = 0, 1, 2

    def measureImageFocusLocal(self, image, avg=True):
        return self._focus(image, self.FOCUS_LOCAL_STD, avg)

    def measureImageFocus(self, image):
        return self._focus(image, self.FOCUS_LOCAL_STD, True)

    def measureImageFocusGlobal(self, image):
        return self._focus(image, self.FOCUS_GLOBAL_STD, True)

    def imageMeanLocal(self, image, avg=True):
        return self._focus(image, self.FOCUS_LOCAL_MEAN, avg)

    def _focus(self, image, what, avg):
        image = np.ascontiguousarray(image, np.float32)
        out = np.zeros(1, np.float32)
        _check(lib.eorb_ev_image_focus(self.h, _p(image), image.shape[1], image.shape[0], image.strides[0], what, 1 if avg else 0, _p(out)),
               "image_focus")
        return float(out[0])

    def image_focus_device(self, d_img_f32, nwin, w, h, what=0, avg=True):
        """metric of nwin device-resident frames -> float32[nwin]; argmax = the reference's best-of-N candidate choice"""
        out = np.zeros(nwin, np.float32)
        _check(lib.eorb_ev_image_focus_device(self.h, _p(d_img_f32), nwin, w, h, what, 1 if avg else 0, _p(out)), "image_focus_device")
        return out

    def ev2im_gauss(self, vEvData, imWidth, imHeight, sigma, pol=False, normalized=True, both=False):
        p = self.make_params(EV_GAUSS, imWidth, imHeight, sigma, pol, NORM_RUNNING if normalized else NORM_NONE)
        _, img, u8, _ = self._run(vEvData, p)
        return (img, u8) if both else (u8 if normalized else img)

    def ev2mci_gg_f(self, vEvData, camera, Tcw, medDepth, imWidth, imHeight, imSigma, pol=False, normalized=True, both=False,
                    norm_mode=None):
        nm = (NORM_RUNNING if normalized else NORM_NONE) if norm_mode is None else norm_mode
        p = self.make_params(EV_SE3, imWidth, imHeight, imSigma, pol, nm, Tcw=Tcw, medDepth=medDepth, camera=camera)
        _, img, u8, _ = self._run(vEvData, p)
        return (img, u8) if both else (u8 if nm != NORM_NONE else img)

    def ev2mci_gg_f_2d(self, vEvData, camera, params2D, imWidth, imHeight, sigma, pol=False, normalized=True, both=False):
        p = self.make_params(EV_SE2, imWidth, imHeight, sigma, pol, NORM_RUNNING if normalized else NORM_NONE, camera=camera,
                             params2D=params2D)
        _, img, u8, _ = self._run(vEvData, p)
        return (img, u8) if both else (u8 if normalized else img)

    def accumulate_batch_device(self, d_evs, win_offsets, p, d_img_f32, d_img_u8=None, poses=None):
        offs = np.ascontiguousarray(win_offsets, np.int64)
        ps = np.ascontiguousarray(poses, np.float32) if poses is not None else None
        _check(lib.eorb_ev_accumulate_batch_device(self.h, _p(d_evs), _p(offs), len(offs) - 1, C.byref(p), _p(ps), _p(d_img_f32),
                                                   _p(d_img_u8)), "ev_accumulate_batch_device")


    def accumulate_batch(self, evs, win_offsets, p, img_f32=None, img_u8=None, poses=None):
        """host buffers: evs = EVENT_DTYPE array (or an int address of pinned memory), outputs = numpy arrays / int addresses [nwin][h][w]"""
        offs = np.ascontiguousarray(win_offsets, np.int64)
        ps = np.ascontiguousarray(poses, np.float32) if poses is not None else None
        _check(lib.eorb_ev_accumulate_batch(self.h, _p(evs), _p(offs), len(offs) - 1, C.byref(p), _p(ps), _p(img_f32), _p(img_u8)),
               "ev_accumulate_batch")


class ELK_Tracker:
    """Mirror of EORB_SLAM::ELK_Tracker (include/Event/KLT_Tracker.h; src/Event/KLT_Tracker.cpp:14-91): setRefImage keeps the
    reference frame and its points, trackCurrImage runs the pyramidal LK of cv::calcOpticalFlowPyrLK against a new frame.
    Defaults are the EvETHZ.yaml values (kltWinSize 23, maxLevel 1, kltMaxItr 10, kltEps 0.03)."""

    def __init__(self, kltWinSize=23, maxLevel=1, kltMaxItr=10, kltEps=0.03, device=0, max_size=(752, 480), max_points=4096):
        self.mPatchSz, self.mMaxLevel, self.maxItr, self.eps = kltWinSize, maxLevel, kltMaxItr, kltEps
        h = C.c_void_p()
        _check(lib.eorb_lk_create(device, max_size[0], max_size[1], max_points, C.byref(h)), "lk_create")
        self.h = h
        self.n = 0

    def __del__(self):
        try:                                   # at interpreter exit the module globals may already be gone
            if getattr(self, "h", None):
                lib.eorb_lk_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def launch_count(self): return lib.eorb_lk_launch_count(self.h)

    def setRefImage(self, image, refPts):
        """refPts: (n, 2) float32 (x, y) — or a KEYPOINT_DTYPE array, whose pt is used (KLT_Tracker.cpp:34-43)"""
        image = np.ascontiguousarray(image, np.uint8)
        if isinstance(refPts, np.ndarray) and refPts.dtype.names:
            refPts = np.stack([refPts["x"], refPts["y"]], 1)
        pts = np.ascontiguousarray(refPts, np.float32).reshape(-1, 2)
        rc = _check(lib.eorb_lk_set_ref(self.h, _p(image), image.shape[1], image.shape[0], image.strides[0], _p(pts), len(pts),
                                        self.mPatchSz, self.mMaxLevel), "lk_set_ref")
        self.n = len(pts) if rc == 0 else 0
        self.shape = image.shape
        return rc

    def trackCurrImage(self, currImage, initPts=None):
        """-> (kpts float32[n,2], status uint8[n], err float32[n]); initPts given = OPTFLOW_USE_INITIAL_FLOW"""
        currImage = np.ascontiguousarray(currImage, np.uint8)
        assert currImage.shape == self.shape
        out = np.zeros((self.n, 2), np.float32); status = np.zeros(self.n, np.uint8); err = np.zeros(self.n, np.float32)
        init = None
        if initPts is not None:
            if isinstance(initPts, np.ndarray) and initPts.dtype.names:
                initPts = np.stack([initPts["x"], initPts["y"]], 1)
            init = np.ascontiguousarray(initPts, np.float32).reshape(-1, 2)
            assert len(init) == self.n
        self.levels_used = _check(lib.eorb_lk_track(self.h, _p(currImage), currImage.strides[0], _p(init), self.maxItr, float(self.eps), 1e-4,
                                                    _p(out), _p(status), _p(err)), "lk_track")
        return out, status, err

    # ---- the tracker with its state on the device (KLT_Tracker.cpp:22-46, 99-262)
    def setRefImageKPts(self, image, refKPts):
        """setRefImage(image, vector<KeyPoint>) :22-46: reference keypoints, reference points and last tracked points (= the points)
        go to the device and stay there"""
        image = np.ascontiguousarray(image, np.uint8)
        kps = np.ascontiguousarray(refKPts, KEYPOINT_DTYPE)
        rc = _check(lib.eorb_lk_set_ref_keypoints(self.h, _p(image), image.shape[1], image.shape[0], image.strides[0], 0, _p(kps), len(kps),
                                                  self.mPatchSz, self.mMaxLevel), "lk_set_ref_keypoints")
        self.n = len(kps) if rc == 0 else 0
        self.shape = image.shape
        return rc

    def setLastTrackedPts(self, currTrackedPts):
        kps = np.ascontiguousarray(currTrackedPts, KEYPOINT_DTYPE)
        return _check(lib.eorb_lk_set_last_tracked(self.h, _p(kps), len(kps)), "lk_set_last_tracked")

    def getLastTrackedPts(self):
        out = np.zeros((self.n, 2), np.float32)
        _check(lib.eorb_lk_get_last_tracked(self.h, _p(out)), "lk_get_last_tracked")
        return out

    def trackAndMatchCurrImage_device(self, d_image, stride, d_tracked, d_matched, d_px_disp, d_counts2, init=False):
        """the same with the frame and every output resident on the device (raw pointers); asynchronous on the tracker's stream.
        Outputs: tracked KEYPOINT_DTYPE[n], matched uint8[n] (bit 0 passed, bit 1 kept), px_disp float32[n], counts int32[2]."""
        return _check(lib.eorb_lk_track_and_match_device(self.h, C.c_void_p(int(d_image)), stride, self.maxItr, float(self.eps), 1e-4,
                                                         1 if init else 0, C.c_void_p(int(d_tracked)), C.c_void_p(int(d_matched)),
                                                         C.c_void_p(int(d_px_disp)), C.c_void_p(int(d_counts2))), "lk_track_and_match_device")

    def set_stream(self, stream):
        if stream is None:
            _check(lib.eorb_lk_reset_stream(self.h), "lk_reset_stream")
        else:
            _check(lib.eorb_lk_set_stream(self.h, C.c_void_p(int(stream))), "lk_set_stream")

    def trackAndMatchCurrImage(self, image, vMatches12=None, vCntMatches=None, init=False):
        """trackAndMatchCurrImage :215-234 (init=True: trackAndMatchCurrImageInit :236-242)
        -> (nMatches, trackedKPts, vMatches12, vCntMatches, vPxDisp).  vMatches12 / vCntMatches: the caller's vectors, updated the way
        the reference updates them in place (None = empty vectors, which the reference resizes to -1 / 1)."""
        image = np.ascontiguousarray(image, np.uint8)
        assert image.shape == self.shape
        n = self.n
        tr = np.zeros(n, KEYPOINT_DTYPE); matched = np.zeros(n, np.uint8); disp = np.zeros(n, np.float32); c2 = np.zeros(2, np.int32)
        rc = _check(lib.eorb_lk_track_and_match(self.h, _p(image), image.strides[0], 0, self.maxItr, float(self.eps), 1e-4, 1 if init else 0,
                                                _p(tr), _p(matched), _p(disp), _p(c2)), "lk_track_and_match")
        m12 = np.full(n, -1, np.int32) if vMatches12 is None else np.array(vMatches12, np.int32)
        cnt = np.ones(n, np.int32) if vCntMatches is None else np.array(vCntMatches, np.int32)
        if rc == EORB_EMPTY:
            return 0, tr[:0], m12, cnt, disp[:0]
        self.levels_used = rc
        passed = (matched & 1).astype(bool)
        idx = np.arange(n, dtype=np.int32)
        # refineTrackedPts :140-145 on the caller's vectors
        m12 = np.where(passed, idx, m12).astype(np.int32); cnt = (cnt + passed).astype(np.int32); nm = int(passed.sum())
        if init:   # refineFirstOctaveLevel :166-176: every entry 0 <= vMatches12[i] < n on an upper reference octave, stale ones included
            hit = (m12 >= 0) & (m12 < n) & (tr["octave"] > 0)
            m12[hit] = -1; cnt[hit] -= 1; nm = (nm - int(hit.sum())) & 0xffffffff   # nMatches is unsigned in the reference
        if vMatches12 is None:
            assert nm == int(c2[0])
        return nm, tr, m12, cnt, disp[:c2[1]].copy()


# ================================================================================================ guided matching
AREA_QUERY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("r", "<f4"), ("min_level", "<i4"), ("max_level", "<i4")])
FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48


class FrameGrid:
    """Mirror of the keypoint grid of ORB_SLAM3::Frame (src/Frame.cc): AssignFeaturesToGrid (:431-460), PosInGrid (:783-793)
    and GetFeaturesInArea (:709-777) for one frame's (undistorted) keypoints.  bounds = (mnMinX, mnMinY, mnMaxX, mnMaxY)."""

    def __init__(self, keypoints, bounds, device=0, guided=None):
        self.kps = np.ascontiguousarray(keypoints, KEYPOINT_DTYPE)
        self.bounds = np.ascontiguousarray(bounds, np.float32)
        self.g = guided or GuidedMatcher(device)

    def AssignFeaturesToGrid(self):
        """-> (cell_start[64*48+1], cell_idx): CSR form of mGrid[col][row], cell id = col*48 + row"""
        cs = np.zeros(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, np.int32); ci = np.zeros(max(len(self.kps), 1), np.int32)
        na = C.c_int(0)
        _check(lib.eorb_guided_frame_grid(self.g.h, _p(self.kps), len(self.kps), _p(self.bounds), _p(cs), _p(ci), C.byref(na)), "frame_grid")
        return cs, ci[:na.value].copy()

    def GetFeaturesInArea(self, x, y, r, minLevel=-1, maxLevel=-1):
        return self.GetFeaturesInAreaBatch(np.array([(x, y, r, minLevel, maxLevel)], AREA_QUERY_DTYPE))[0]

    def GetFeaturesInAreaBatch(self, queries):
        """queries: AREA_QUERY_DTYPE array -> list of index arrays in the reference's order"""
        q = np.ascontiguousarray(queries, AREA_QUERY_DTYPE)
        cap = max(len(self.kps), 1)
        cnt = np.zeros(max(len(q), 1), np.int32); out = np.zeros((max(len(q), 1), cap), np.int32)
        _check(lib.eorb_guided_features_in_area(self.g.h, _p(self.kps), len(self.kps), _p(self.bounds), _p(q), len(q), _p(cnt), _p(out), cap),
               "features_in_area")
        return [out[k, :cnt[k]].copy() for k in range(len(q))]


class GuidedMatcher:
    """Mirror of the guided-matching entry points of ORB_SLAM3::ORBmatcher (src/ORBmatcher.cc) that walk the Frame grid:
    SearchForInitialization (:714-831)."""

    def __init__(self, device=0, nnratio=0.9, checkOri=True):
        h = C.c_void_p()
        _check(lib.eorb_guided_create(device, C.byref(h)), "guided_create")
        self.h = h
        self.mfNNratio = float(nnratio); self.mbCheckOrientation = bool(checkOri)

    def __del__(self):
        try:                                   # at interpreter exit the module globals may already be gone
            if getattr(self, "h", None):
                lib.eorb_guided_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_stream(self, s):
        _check(lib.eorb_guided_set_stream(self.h, C.c_void_p(s)) if s is not None else lib.eorb_guided_reset_stream(self.h), "guided_set_stream")

    def launch_count(self): return lib.eorb_guided_launch_count(self.h)

    def SearchForInitialization(self, kps1, desc1, kps2, desc2, bounds, vbPrevMatched, windowSize=100):
        """-> (nmatches, vnMatches12[n1], vbPrevMatched updated copy (n1, 2))"""
        k1 = np.ascontiguousarray(kps1, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
        d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        b = np.ascontiguousarray(bounds, np.float32)
        prev = np.array(vbPrevMatched, np.float32, copy=True).reshape(-1, 2)
        m12 = np.full(max(len(k1), 1), -1, np.int32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_for_initialization(self.h, _p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2), _p(b), _p(prev),
                                                         int(windowSize), self.mfNNratio, int(self.mbCheckOrientation), _p(m12),
                                                         C.byref(nm)), "SearchForInitialization")
        return nm.value, m12[:len(k1)].copy(), prev

    def SearchByProjection(self, x3Dc, valid1, obs1, kps1, descMP, kps2, desc2, bounds, K4, scale_factors, th=15.0):
        """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono=True) (:1969-2150) -> (nmatches, match_cur[n2]):
        match_cur[i2] = last-frame index whose map point is assigned to current-frame keypoint i2, or -1"""
        x = np.ascontiguousarray(x3Dc, np.float32).reshape(-1, 3); v = np.ascontiguousarray(valid1, np.uint8); o = np.ascontiguousarray(obs1, np.int32)
        k1 = np.ascontiguousarray(kps1, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
        dm = np.ascontiguousarray(descMP, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        b = np.ascontiguousarray(bounds, np.float32); K = np.ascontiguousarray(K4, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
        mc = np.full(max(len(k2), 1), -1, np.int32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_by_projection(self.h, _p(x), _p(v), _p(o), _p(k1), _p(dm), len(k1), _p(k2), _p(d2), len(k2), _p(b), _p(K),
                                                    _p(sf), len(sf), float(th), int(self.mbCheckOrientation), _p(mc), C.byref(nm)),
               "SearchByProjection")
        return nm.value, mc[:len(k2)].copy()

    def SearchByProjectionStereo(self, x3Dc, valid1, obs1, kps1, descMP, kps2, desc2, bounds, K4, scale_factors, th=15.0, level_mode=0,
                                 mbf=0.0, u_right2=None):
        """SearchByProjection(CurrentFrame, LastFrame, th, bMono=false) for a rectified-stereo / RGB-D frame (:1989-1990 forward /
        backward level windows, :2049-2055 right-column test) -> (nmatches, match_cur[n2])"""
        x = np.ascontiguousarray(x3Dc, np.float32).reshape(-1, 3); v = np.ascontiguousarray(valid1, np.uint8); o = np.ascontiguousarray(obs1, np.int32)
        k1 = np.ascontiguousarray(kps1, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
        dm = np.ascontiguousarray(descMP, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        b = np.ascontiguousarray(bounds, np.float32); K = np.ascontiguousarray(K4, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
        ur = None if u_right2 is None else np.ascontiguousarray(u_right2, np.float32)
        mc = np.full(max(len(k2), 1), -1, np.int32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_by_projection_stereo(self.h, _p(x), _p(v), _p(o), _p(k1), _p(dm), len(k1), _p(k2), _p(d2),
                                                           _p(ur) if ur is not None else None, len(k2), _p(b), _p(K), _p(sf), len(sf), float(th),
                                                           int(self.mbCheckOrientation), int(level_mode), float(mbf), _p(mc), C.byref(nm)),
               "SearchByProjection(stereo)")
        return nm.value, mc[:len(k2)].copy()

    def SearchByProjectionReloc(self, x3Dc, valid1, level1, kps1, descMP, kps2, desc2, held2, bounds, K4, scale_factors, th=10.0, ORBdist=100):
        """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (:2189-2312; Tracking::Relocalization)
        -> (nmatches, match_cur[n2])"""
        x = np.ascontiguousarray(x3Dc, np.float32).reshape(-1, 3); v = np.ascontiguousarray(valid1, np.uint8); lv = np.ascontiguousarray(level1, np.int32)
        k1 = np.ascontiguousarray(kps1, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
        dm = np.ascontiguousarray(descMP, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        b = np.ascontiguousarray(bounds, np.float32); K = np.ascontiguousarray(K4, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
        hd = None if held2 is None else np.ascontiguousarray(held2, np.uint8)
        mc = np.full(max(len(k2), 1), -1, np.int32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_by_projection_reloc(self.h, _p(x), _p(v), _p(lv), _p(k1), _p(dm), len(k1), _p(k2), _p(d2),
                                                          _p(hd) if hd is not None else None, len(k2), _p(b), _p(K), _p(sf), len(sf), float(th),
                                                          int(ORBdist), int(self.mbCheckOrientation), _p(mc), C.byref(nm)),
               "SearchByProjection(relocalisation)")
        return nm.value, mc[:len(k2)].copy()

    def SearchByProjectionMapPointsStereo(self, pts, proj_xr, descMP, kps2, desc2, held2, u_right2, bounds, scale_factors, th=1.0, bFarPoints=False,
                                          thFarPoints=0.0):
        """SearchByProjection(F, vpMapPoints, ...) for a rectified-stereo / RGB-D frame (:91-96) -> (nmatches, match_cur[n2])"""
        p = np.ascontiguousarray(pts, TRACK_POINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
        dm = np.ascontiguousarray(descMP, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        hd = None if held2 is None else np.ascontiguousarray(held2, np.uint8)
        xr = None if proj_xr is None else np.ascontiguousarray(proj_xr, np.float32)
        ur = None if u_right2 is None else np.ascontiguousarray(u_right2, np.float32)
        b = np.ascontiguousarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
        mc = np.full(max(len(k2), 1), -1, np.int32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_by_projection_map_points_stereo(self.h, _p(p), _p(xr) if xr is not None else None, _p(dm), len(p), _p(k2), _p(d2),
                                                                      _p(hd) if hd is not None else None, _p(ur) if ur is not None else None,
                                                                      len(k2), _p(b), _p(sf), len(sf), C.c_float(th), int(bool(bFarPoints)),
                                                                      C.c_float(thFarPoints), C.c_float(self.mfNNratio), _p(mc), C.byref(nm)),
               "SearchByProjection(map points, stereo)")
        return nm.value, mc[:len(k2)].copy()

    def SearchByProjectionMapPoints(self, pts, descMP, kps2, desc2, held2, bounds, scale_factors, th=1.0, bFarPoints=False, thFarPoints=0.0):
        """ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) (:44-218, monocular frame;
        Tracking::SearchLocalPoints) -> (nmatches, match_cur[n2]).  pts: TRACK_POINT_DTYPE (what Frame::isInFrustum fills)."""
        p = np.ascontiguousarray(pts, TRACK_POINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
        dm = np.ascontiguousarray(descMP, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        hd = None if held2 is None else np.ascontiguousarray(held2, np.uint8)
        b = np.ascontiguousarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
        mc = np.full(max(len(k2), 1), -1, np.int32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_by_projection_map_points(self.h, _p(p), _p(dm), len(p), _p(k2), _p(d2), _p(hd) if hd is not None else None,
                                                               len(k2), _p(b), _p(sf), len(sf), C.c_float(th), int(bool(bFarPoints)),
                                                               C.c_float(thFarPoints), C.c_float(self.mfNNratio), _p(mc), C.byref(nm)),
               "SearchByProjection(map points)")
        return nm.value, mc[:len(k2)].copy()

    def SearchForTriangulation(self, kps1, desc1, flags1, fv1, kps2, desc2, flags2, fv2, F12, epipole2, scale_factors2, level_sigma2_2, bCoarse=False):
        """ORBmatcher::SearchForTriangulation (:975-1214), pinhole keyframes -> (nmatches, match12[n1]); flags: bit 0 = takes part, bit 1 = bStereo"""
        k1 = np.ascontiguousarray(kps1, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
        d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        f1 = np.ascontiguousarray(flags1, np.uint8); f2 = np.ascontiguousarray(flags2, np.uint8)
        a = [np.ascontiguousarray(fv1[0], np.uint32), np.ascontiguousarray(fv1[1], np.int32), np.ascontiguousarray(fv1[2], np.uint32)]
        b = [np.ascontiguousarray(fv2[0], np.uint32), np.ascontiguousarray(fv2[1], np.int32), np.ascontiguousarray(fv2[2], np.uint32)]
        F = np.ascontiguousarray(F12, np.float32).reshape(9); e = np.ascontiguousarray(epipole2, np.float32)
        sc = np.ascontiguousarray(scale_factors2, np.float32); sg = np.ascontiguousarray(level_sigma2_2, np.float32)
        m12 = np.full(max(len(k1), 1), -1, np.int32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_for_triangulation(self.h, _p(k1), _p(d1), _p(f1), len(k1), _p(a[0]), _p(a[1]), _p(a[2]), len(a[0]), _p(k2), _p(d2),
                                                        _p(f2), len(k2), _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]), _p(F), _p(e), _p(sc), _p(sg), len(sc),
                                                        int(bCoarse), int(self.mbCheckOrientation), _p(m12), C.byref(nm)),
               "SearchForTriangulation")
        return nm.value, m12[:len(k1)].copy()

    def SearchForTriangulation_device(self, d_kps1, d_desc1, d_flags1, n1, d_fv1, nn1, nentries1, d_kps2, d_desc2, d_flags2, n2, d_fv2, nn2, F12, epipole2,
                                      scale_factors2, level_sigma2_2, d_match12, bCoarse=False):
        """device pointers (ints); d_fv = (nodes, start, feats) pointers; returns nmatches"""
        F = np.ascontiguousarray(F12, np.float32).reshape(9); e = np.ascontiguousarray(epipole2, np.float32)
        sc = np.ascontiguousarray(scale_factors2, np.float32); sg = np.ascontiguousarray(level_sigma2_2, np.float32)
        nm = C.c_int(0)
        vp = C.c_void_p
        _check(lib.eorb_guided_search_for_triangulation_device(self.h, vp(d_kps1), vp(d_desc1), vp(d_flags1), n1, vp(d_fv1[0]), vp(d_fv1[1]), vp(d_fv1[2]), nn1,
                                                               nentries1, vp(d_kps2), vp(d_desc2), vp(d_flags2), n2, vp(d_fv2[0]), vp(d_fv2[1]), vp(d_fv2[2]),
                                                               nn2, _p(F), _p(e), _p(sc), _p(sg), len(sc), int(bCoarse), int(self.mbCheckOrientation),
                                                               vp(d_match12), C.byref(nm)), "SearchForTriangulation_device")
        return nm.value

    def SearchWindows(self, queries, ur, descMP, kps2, desc2, held2, u_right2, bounds, query_min_xy=None, inv_level_sigma2=None, blocking=False,
                      th_high=50):
        """eorb_guided_search_windows: the matching core of the keyframe-side searches -- ORBmatcher::SearchByProjection(KeyFrame*, Scw, ...)
        (:480, :595), Fuse (:1407, :1619), SearchBySim3 (:1743) -> (nmatches, best_idx[n1], best_dist[n1], match2[n2])"""
        q = np.ascontiguousarray(queries, AREA_QUERY_DTYPE)
        k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE); dm = np.ascontiguousarray(descMP, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        b = np.ascontiguousarray(bounds, np.float32)
        opt = lambda a, t: None if a is None else np.ascontiguousarray(a, t)
        urq, hd, ur2, qm, inv = opt(ur, np.float32), opt(held2, np.uint8), opt(u_right2, np.float32), opt(query_min_xy, np.float32), opt(inv_level_sigma2, np.float32)
        pp = lambda a: _p(a) if a is not None else None
        bi = np.full(max(len(q), 1), -1, np.int32); bd = np.full(max(len(q), 1), 256, np.int32); m2 = np.full(max(len(k2), 1), -1, np.int32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_windows(self.h, _p(q), pp(urq), _p(dm), len(q), _p(k2), _p(d2), pp(hd), pp(ur2), len(k2), _p(b), pp(qm), pp(inv),
                                              0 if inv is None else len(inv), int(blocking), int(th_high), _p(bi), _p(bd), _p(m2), C.byref(nm)),
               "SearchWindows")
        return nm.value, bi[:len(q)].copy(), bd[:len(q)].copy(), m2[:len(k2)].copy()

    def SearchWindows_device(self, d_queries, d_ur, d_descMP, n1, d_kps2, d_desc2, d_held2, d_u_right2, n2, bounds, d_best_idx, d_best_dist,
                             d_match2, query_min_xy=None, inv_level_sigma2=None, blocking=False, th_high=50):
        """device pointers (ints; optional ones may be 0); returns nmatches"""
        b = np.ascontiguousarray(bounds, np.float32)
        qm = None if query_min_xy is None else np.ascontiguousarray(query_min_xy, np.float32)
        inv = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
        nm = C.c_int(0)
        vp = lambda a: C.c_void_p(a) if a else None
        _check(lib.eorb_guided_search_windows_device(self.h, vp(d_queries), vp(d_ur), vp(d_descMP), n1, vp(d_kps2), vp(d_desc2), vp(d_held2),
                                                     vp(d_u_right2), n2, _p(b), _p(qm) if qm is not None else None,
                                                     _p(inv) if inv is not None else None, 0 if inv is None else len(inv), int(blocking),
                                                     int(th_high), vp(d_best_idx), vp(d_best_dist), vp(d_match2), C.byref(nm)),
               "SearchWindows_device")
        return nm.value

    def SearchByBoW(self, kpsKF, descKF, validKF, fvKF, kpsF, descF, fvF):
        """ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (:276-478, monocular) -> (nmatches, match_f[n2]); fv* = (nodes, start,
        feats): the FeatureVector in the CSR form ORBVocabulary.transform returns (fv_nodes, fv_start, fv_feats)"""
        k1 = np.ascontiguousarray(kpsKF, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kpsF, KEYPOINT_DTYPE)
        d1 = np.ascontiguousarray(descKF, np.uint8); d2 = np.ascontiguousarray(descF, np.uint8); v = np.ascontiguousarray(validKF, np.uint8)
        a = [np.ascontiguousarray(fvKF[0], np.uint32), np.ascontiguousarray(fvKF[1], np.int32), np.ascontiguousarray(fvKF[2], np.uint32)]
        b = [np.ascontiguousarray(fvF[0], np.uint32), np.ascontiguousarray(fvF[1], np.int32), np.ascontiguousarray(fvF[2], np.uint32)]
        mf = np.full(max(len(k2), 1), -1, np.int32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_by_bow(self.h, _p(k1), _p(d1), _p(v), len(k1), _p(a[0]), _p(a[1]), _p(a[2]), len(a[0]), _p(k2), _p(d2), len(k2),
                                             _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]), C.c_float(self.mfNNratio), int(self.mbCheckOrientation), _p(mf),
                                             C.byref(nm)), "SearchByBoW")
        return nm.value, mf[:len(k2)].copy()

    def SearchByBoW_KF(self, kps1, desc1, valid1, fv1, kps2, desc2, valid2, fv2):
        """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (:833-990, monocular keyframes) -> (nmatches, match12[n1]): match12[i1] = feature
        index of keyframe 2 whose map point the reference stores in vpMatches12[i1], or -1; valid* = a map point that is not bad"""
        k1 = np.ascontiguousarray(kps1, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(kps2, KEYPOINT_DTYPE)
        d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        v1 = np.ascontiguousarray(valid1, np.uint8); v2 = np.ascontiguousarray(valid2, np.uint8)
        a = [np.ascontiguousarray(fv1[0], np.uint32), np.ascontiguousarray(fv1[1], np.int32), np.ascontiguousarray(fv1[2], np.uint32)]
        b = [np.ascontiguousarray(fv2[0], np.uint32), np.ascontiguousarray(fv2[1], np.int32), np.ascontiguousarray(fv2[2], np.uint32)]
        m12 = np.full(max(len(k1), 1), -1, np.int32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_by_bow_kf(self.h, _p(k1), _p(d1), _p(v1), len(k1), _p(a[0]), _p(a[1]), _p(a[2]), len(a[0]), _p(k2), _p(d2), _p(v2),
                                                len(k2), _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]), C.c_float(self.mfNNratio),
                                                int(self.mbCheckOrientation), _p(m12), C.byref(nm)), "SearchByBoW_KF")
        return nm.value, m12[:len(k1)].copy()

    def SearchByBoW_device(self, d_kpsKF, d_descKF, d_validKF, n1, d_fvKF, nkf, d_kpsF, d_descF, n2, d_fvF, nf, d_match_f):
        """device pointers (ints); d_fv* = (nodes, start, feats) device pointers; returns nmatches"""
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_by_bow_device(self.h, C.c_void_p(d_kpsKF), C.c_void_p(d_descKF), C.c_void_p(d_validKF), n1,
                                                    C.c_void_p(d_fvKF[0]), C.c_void_p(d_fvKF[1]), C.c_void_p(d_fvKF[2]), nkf, C.c_void_p(d_kpsF),
                                                    C.c_void_p(d_descF), n2, C.c_void_p(d_fvF[0]), C.c_void_p(d_fvF[1]), C.c_void_p(d_fvF[2]), nf,
                                                    C.c_float(self.mfNNratio), int(self.mbCheckOrientation), C.c_void_p(d_match_f), C.byref(nm)),
               "SearchByBoW_device")
        return nm.value

    def SearchByProjectionMapPoints_device(self, d_pts, d_descMP, n1, d_kps2, d_desc2, d_held2, n2, bounds, scale_factors, d_match_cur, th=1.0,
                                           bFarPoints=False, thFarPoints=0.0):
        """device pointers (ints; d_held2 may be 0); returns nmatches"""
        b = np.ascontiguousarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_by_projection_map_points_device(self.h, C.c_void_p(d_pts), C.c_void_p(d_descMP), n1, C.c_void_p(d_kps2),
                                                                      C.c_void_p(d_desc2), C.c_void_p(d_held2) if d_held2 else None, n2, _p(b), _p(sf),
                                                                      len(sf), C.c_float(th), int(bool(bFarPoints)), C.c_float(thFarPoints),
                                                                      C.c_float(self.mfNNratio), C.c_void_p(d_match_cur), C.byref(nm)),
               "SearchByProjectionMapPoints_device")
        return nm.value

    def SearchForInitialization_device(self, d_kps1, d_desc1, n1, d_kps2, d_desc2, n2, bounds, d_prev, d_matches12, windowSize=100):
        """all pointers are device addresses (ints); returns nmatches"""
        b = np.ascontiguousarray(bounds, np.float32)
        nm = C.c_int(0)
        _check(lib.eorb_guided_search_for_initialization_device(self.h, C.c_void_p(d_kps1), C.c_void_p(d_desc1), n1, C.c_void_p(d_kps2),
                                                                C.c_void_p(d_desc2), n2, _p(b), C.c_void_p(d_prev), int(windowSize),
                                                                self.mfNNratio, int(self.mbCheckOrientation), C.c_void_p(d_matches12),
                                                                C.byref(nm)), "SearchForInitialization_device")
        return nm.value


# ================================================================================================ bag of words + undistortion
class ORBVocabulary:
    """Mirror of ORB_SLAM3::ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> for the call the front end
    makes: transform(features, BowVector, FeatureVector, levelsup) (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1258;
    Frame::ComputeBoW src/Frame.cc:796-803).  `voc` is the flat tree (see eorb_vocab_create / synth.make_vocabulary)."""

    def __init__(self, voc, device=0):
        h = C.c_void_p()
        par = np.ascontiguousarray(voc["parent"], np.int32); lf = np.ascontiguousarray(voc["is_leaf"], np.uint8)
        d = np.ascontiguousarray(voc["desc"], np.uint8); w = np.ascontiguousarray(voc["weight"], np.float64)
        _check(lib.eorb_vocab_create(device, int(voc["k"]), int(voc["L"]), int(voc["scoring"]), int(voc["weighting"]), len(par), _p(par), _p(lf),
                                     _p(d), _p(w), C.byref(h)), "vocab_create")
        self.h = h

    def __del__(self):
        try:                                   # at interpreter exit the module globals may already be gone
            if getattr(self, "h", None):
                lib.eorb_vocab_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_stream(self, s):
        _check(lib.eorb_vocab_set_stream(self.h, C.c_void_p(s)) if s is not None else lib.eorb_vocab_reset_stream(self.h), "vocab_set_stream")

    def launch_count(self): return lib.eorb_vocab_launch_count(self.h)

    def _run(self, fn, feats_ptr, n, levelsup):
        m = max(n, 1)
        bi = np.zeros(m, np.uint32); bv = np.zeros(m, np.float64); fn_ = np.zeros(m, np.uint32); fs = np.zeros(m + 1, np.int32)
        ff = np.zeros(m, np.uint32); wid = np.zeros(m, np.uint32); nid = np.zeros(m, np.uint32)
        nb = C.c_int(0); nf = C.c_int(0)
        _check(fn(self.h, feats_ptr, n, int(levelsup), _p(bi), _p(bv), C.byref(nb), _p(fn_), _p(fs), _p(ff), C.byref(nf), _p(wid), _p(nid)),
               "vocab_transform")
        nbow, nfv = nb.value, nf.value
        return dict(word_id=wid[:n], node_id=nid[:n], bow_ids=bi[:nbow].copy(), bow_vals=bv[:nbow].copy(), fv_nodes=fn_[:nfv].copy(),
                    fv_start=fs[:nfv + 1].copy(), fv_feats=ff[:fs[nfv]].copy())

    def transform(self, features, levelsup=4):
        """features: (n, 32) u8 -> dict(bow_ids, bow_vals (BowVector), fv_nodes, fv_start, fv_feats (FeatureVector), word_id, node_id)"""
        f = np.ascontiguousarray(features, np.uint8).reshape(-1, 32)
        return self._run(lib.eorb_vocab_transform, _p(f), len(f), levelsup)

    def transform_resident(self, d_feats_ptr, n, levelsup, d_fv_nodes, d_fv_start, d_fv_feats, d_bow_ids=0, d_bow_vals=0):
        """descriptors AND results in HBM (device addresses as ints) -> (nbow, nfv)"""
        nb, nf = C.c_int(0), C.c_int(0)
        _check(lib.eorb_vocab_transform_resident(self.h, C.c_void_p(d_feats_ptr), int(n), int(levelsup), C.c_void_p(d_bow_ids) if d_bow_ids else None,
                                                 C.c_void_p(d_bow_vals) if d_bow_vals else None, C.byref(nb), C.c_void_p(d_fv_nodes),
                                                 C.c_void_p(d_fv_start), C.c_void_p(d_fv_feats), C.byref(nf)), "transform_resident")
        return nb.value, nf.value

    def transform_device(self, d_feats_ptr, n, levelsup=4):
        return self._run(lib.eorb_vocab_transform_device, C.c_void_p(d_feats_ptr), n, levelsup)


def UndistortKeyPoints_device(d_in, d_out, n, K4, distCoef5, stream=None):
    """Frame::UndistortKeyPoints on device arrays (ints = device addresses; d_out may equal d_in)"""
    K = np.ascontiguousarray(K4, np.float32); D = np.ascontiguousarray(distCoef5, np.float32)
    _check(lib.eorb_undistort_keypoints_device(C.c_void_p(d_in), C.c_void_p(d_out), int(n), _p(K), _p(D), C.c_void_p(stream) if stream else None),
           "UndistortKeyPoints_device")


def UndistortKeyPoints(keypoints, K4, distCoef5):
    """Mirror of Frame::UndistortKeyPoints (src/Frame.cc:805-840): -> mvKeysUn (KEYPOINT_DTYPE array)"""
    k = np.ascontiguousarray(keypoints, KEYPOINT_DTYPE)
    out = np.zeros_like(k)
    K = np.ascontiguousarray(K4, np.float32); D = np.ascontiguousarray(distCoef5, np.float32)
    _check(lib.eorb_undistort_keypoints(_p(k), len(k), _p(K), _p(D), _p(out)), "UndistortKeyPoints")
    return out
