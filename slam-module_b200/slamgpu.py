"""ctypes binding of csrc/libslamgpu.so (C ABI: include/slamgpu.h) plus thin host-side mirrors of
the reference interfaces this path replaces:

    ImagePyramid.update / get_level / get_blurred_level   (image_pyramid.hpp:16-30)
    FeatureDetector.detect                                (feature_detector.hpp:15-24)
    OrbExtractor.detect_and_extract                       (orb_extractor.hpp:11-30)
    match_for_loop_closures (brute-force case)            (keyframe_matcher.hpp:33-40)

This is harness code for tests and benches: all work happens behind the C ABI on the GPU.  There
is no CPU fallback -- if the library or a CUDA device is missing, construction raises.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
LIB_PATH = _DIR / "csrc" / "libslamgpu.so"
MAX_LEVELS = 16

SG_OK, SG_ERR_INVALID, SG_ERR_CUDA, SG_ERR_OVERFLOW, SG_ERR_STATE = 0, 1, 2, 3, 4


class SlamGpuError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("libslamgpu error %d: %s" % (code, text))
        self.code = code


class Params(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("levels", C.c_int), ("scale_factor", C.c_float),
                ("max_keypoints", C.c_int), ("ini_fast_thr", C.c_int), ("min_fast_thr", C.c_int),
                ("max_frames", C.c_int), ("max_tracks", C.c_int), ("track_level", C.c_int)]


class MatchParams(C.Structure):
    _fields_ = [("ratio", C.c_float), ("thr", C.c_uint32), ("check_orientation", C.c_int),
                ("ratio_is_double", C.c_int)]


class Keypoints(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("angle", C.c_void_p), ("octave", C.c_void_p),
                ("desc", C.c_void_p), ("track_id", C.c_void_p), ("lvl_x", C.c_void_p), ("lvl_y", C.c_void_p),
                ("count", C.c_void_p), ("level_count", C.c_void_p)]


class TriangulationParams(C.Structure):
    _fields_ = [("E", C.c_double * 9), ("scale_factors", C.c_void_p), ("n_levels", C.c_int), ("residual_deg_thr", C.c_float),
                ("thr", C.c_uint32), ("check_orientation", C.c_int)]


class KeypointsDev(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("angle", C.c_void_p), ("octave", C.c_void_p),
                ("desc", C.c_void_p), ("count", C.c_void_p), ("cap", C.c_int)]


# every symbol include/slamgpu.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "sg_create", "sg_destroy", "sg_last_error", "sg_abi_version", "sg_synchronize", "sg_stream",
    "sg_launch_count", "sg_get_geometry", "sg_keypoint_capacity", "sg_pyramid_update",
    "sg_pyramid_update_device", "sg_pyramid_download", "sg_pyramid_device_plane", "sg_detect",
    "sg_detect_download", "sg_detect_download_candidates", "sg_extract", "sg_extract_device",
    "sg_extract_download", "sg_extract_device_views", "sg_hamming", "sg_match_bruteforce", "sg_db_create",
    "sg_db_create_device", "sg_db_destroy", "sg_match_pairs", "sg_match_pairs_device", "sg_match_rescans",
    "sg_angle_bin_order", "sg_angle_bin_order_depth", "sg_angle_bin", "sg_device_count", "sg_malloc", "sg_free",
    "sg_memcpy_h2d", "sg_memcpy_d2h", "sg_host_alloc_pinned", "sg_host_free_pinned", "sg_timer_start",
    "sg_timer_stop", "sg_flush_l2", "sg_microbench_popc", "sg_set_profiling", "sg_get_stage_ms",
    "sg_set_pipeline_chunk", "sg_search_candidates", "sg_match_sim3", "sg_feature_index", "sg_medoid", "sg_set_overlap", "sg_match_bow", "sg_vocab_create", "sg_vocab_destroy",
    "sg_bow_transform", "sg_bow_transform_device", "sg_match_triangulation", "sg_bow_vector", "sg_bowdb_create",
    "sg_bowdb_destroy", "sg_bowdb_size", "sg_bowdb_add", "sg_bowdb_remove", "sg_bow_similar", "sg_extract_submit",
    "sg_extract_wait", "sg_db_wrap_device", "sg_bow_vector_batch",
]

_lib = None


def lib():
    """Load libslamgpu.so.  Raises if it has not been built: there is no other implementation."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(the CUDA library is the only implementation of this path)" % LIB_PATH)
        L = C.CDLL(str(LIB_PATH))
        L.sg_last_error.restype = C.c_char_p
        L.sg_last_error.argtypes = [C.c_void_p]
        L.sg_stream.restype = C.c_void_p
        L.sg_stream.argtypes = [C.c_void_p]
        L.sg_launch_count.restype = C.c_ulonglong
        L.sg_launch_count.argtypes = [C.c_void_p]
        L.sg_match_rescans.restype = C.c_ulonglong
        L.sg_match_rescans.argtypes = [C.c_void_p]
        L.sg_angle_bin.argtypes = [C.c_float]
        L.sg_destroy.argtypes = [C.c_void_p]
        L.sg_destroy.restype = None
        L.sg_db_destroy.argtypes = [C.c_void_p]
        L.sg_db_destroy.restype = None
        for name in ("sg_create", "sg_synchronize", "sg_get_geometry", "sg_keypoint_capacity", "sg_pyramid_update",
                     "sg_pyramid_update_device", "sg_pyramid_download", "sg_pyramid_device_plane", "sg_detect",
                     "sg_detect_download", "sg_detect_download_candidates", "sg_extract", "sg_extract_device",
                     "sg_extract_download", "sg_extract_device_views", "sg_hamming", "sg_match_bruteforce",
                     "sg_db_create", "sg_db_create_device", "sg_match_pairs", "sg_match_pairs_device", "sg_malloc",
                     "sg_free", "sg_memcpy_h2d", "sg_memcpy_d2h", "sg_timer_start", "sg_timer_stop", "sg_flush_l2",
                     "sg_microbench_popc"):
            getattr(L, name).restype = C.c_int
        L.sg_pyramid_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int]
        L.sg_pyramid_update_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int]
        L.sg_pyramid_download.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.sg_detect_download.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sg_detect_download_candidates.argtypes = L.sg_detect_download.argtypes
        L.sg_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
        L.sg_extract_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int]
        L.sg_extract_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.sg_extract_submit.restype = C.c_int
        L.sg_extract_wait.argtypes = [C.c_void_p, C.c_int]
        L.sg_extract_wait.restype = C.c_int
        L.sg_extract_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.sg_hamming.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sg_match_bruteforce.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
        L.sg_db_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sg_db_create_device.argtypes = L.sg_db_create.argtypes
        L.sg_db_wrap_device.argtypes = L.sg_db_create.argtypes
        L.sg_db_wrap_device.restype = C.c_int
        L.sg_match_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sg_match_pairs_device.argtypes = L.sg_match_pairs.argtypes
        L.sg_malloc.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.sg_free.argtypes = [C.c_void_p, C.c_void_p]
        L.sg_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.sg_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.sg_host_alloc_pinned.argtypes = [C.c_size_t, C.c_void_p]
        L.sg_host_free_pinned.argtypes = [C.c_void_p]
        L.sg_timer_start.argtypes = [C.c_void_p]
        L.sg_timer_stop.argtypes = [C.c_void_p, C.c_void_p]
        L.sg_flush_l2.argtypes = [C.c_void_p]
        L.sg_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.sg_set_profiling.restype = C.c_int
        L.sg_get_stage_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sg_get_stage_ms.restype = C.c_int
        L.sg_synchronize.argtypes = [C.c_void_p]
        L.sg_set_pipeline_chunk.argtypes = [C.c_void_p, C.c_int]
        L.sg_search_candidates.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p] + [C.c_void_p] * 5 + \
            [C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sg_search_candidates.restype = C.c_int
        L.sg_feature_index.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sg_feature_index.restype = C.c_int
        L.sg_medoid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sg_match_sim3.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p] + \
                                   [C.c_void_p] * 10 + [C.c_void_p, C.c_void_p]
        L.sg_match_sim3.restype = C.c_int
        L.sg_medoid.restype = C.c_int
        L.sg_set_pipeline_chunk.restype = C.c_int
        L.sg_set_overlap.argtypes = [C.c_void_p, C.c_int]
        L.sg_set_overlap.restype = C.c_int
        L.sg_match_bow.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3
        L.sg_match_bow.restype = C.c_int
        L.sg_match_triangulation.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 3
        L.sg_match_triangulation.restype = C.c_int
        L.sg_vocab_create.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_void_p]
        L.sg_vocab_create.restype = C.c_int
        L.sg_vocab_destroy.argtypes = [C.c_void_p]
        L.sg_vocab_destroy.restype = None
        L.sg_bow_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sg_bow_transform.restype = C.c_int
        L.sg_bow_transform_device.argtypes = L.sg_bow_transform.argtypes
        L.sg_bow_transform_device.restype = C.c_int
        L.sg_bow_vector.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sg_bow_vector.restype = C.c_int
        L.sg_bow_vector_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sg_bow_vector_batch.restype = C.c_int
        L.sg_bowdb_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.sg_bowdb_create.restype = C.c_int
        L.sg_bowdb_destroy.argtypes = [C.c_void_p]
        L.sg_bowdb_destroy.restype = None
        L.sg_bowdb_size.argtypes = [C.c_void_p]
        L.sg_bowdb_size.restype = C.c_int
        L.sg_bowdb_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.sg_bowdb_add.restype = C.c_int
        L.sg_bowdb_remove.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.sg_bowdb_remove.restype = C.c_int
        L.sg_bow_similar.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sg_bow_similar.restype = C.c_int
        L.sg_detect.argtypes = [C.c_void_p]
        L.sg_keypoint_capacity.argtypes = [C.c_void_p]
        L.sg_microbench_popc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sg_get_geometry.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.sg_extract_device_views.argtypes = [C.c_void_p, C.c_void_p]
        L.sg_pyramid_device_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def device_count():
    return int(lib().sg_device_count())


def feature_index(x, y):
    """FeatureSearch's Y-sorted keypoint order (host code of the library, no GPU)."""
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    order = np.empty(max(len(x), 1), np.int32)
    lib().sg_feature_index(x.ctypes.data, y.ctypes.data, len(x), order.ctypes.data)
    return order[:len(x)]


def angle_bin_order(sizes, depth_limit=None):
    sizes = np.ascontiguousarray(sizes, np.uint32)
    assert sizes.shape == (30,)
    out = np.empty(30, np.uint32)
    if depth_limit is None:
        lib().sg_angle_bin_order(sizes.ctypes, out.ctypes)
    else:
        lib().sg_angle_bin_order_depth(sizes.ctypes, int(depth_limit), out.ctypes)
    return out


def angle_bin(delta):
    return int(lib().sg_angle_bin(C.c_float(delta)))


class DeviceBuffer:
    """A device allocation owned through the C ABI (the harness has no CUDA binding of its own)."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        ctx._check(lib().sg_malloc(ctx._h, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.ctx._check(lib().sg_memcpy_h2d(self.ctx._h, self.ptr, arr.ctypes.data, arr.nbytes))
        return self

    def download(self, dtype, count):
        out = np.empty(count, dtype)
        assert out.nbytes <= self.nbytes
        self.ctx._check(lib().sg_memcpy_d2h(self.ctx._h, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().sg_free(self.ctx._h, self.ptr)
            self.ptr = None


class PinnedArray:
    """numpy view of page-locked host memory (for end-to-end timing with real H2D / D2H copies)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        p = C.c_void_p()
        if lib().sg_host_alloc_pinned(max(self.nbytes, 1), C.byref(p)) != 0:
            raise SlamGpuError(SG_ERR_CUDA, "cudaMallocHost failed")
        self.ptr = p.value
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().sg_host_free_pinned(self.ptr)
            self.ptr = None


class Context:
    """One sg_ctx: a (host thread, GPU) pair sized for a fixed image size and batch."""

    def __init__(self, width, height, levels=8, scale_factor=1.2, max_keypoints=2000, ini_fast_thr=20,
                 min_fast_thr=7, max_frames=1, max_tracks=0, track_level=0, device=0):
        self.params = Params(width, height, levels, scale_factor, max_keypoints, ini_fast_thr, min_fast_thr,
                             max_frames, max_tracks, track_level)
        h = C.c_void_p()
        rc = lib().sg_create(int(device), C.byref(self.params), C.byref(h))
        if rc != 0:
            raise SlamGpuError(rc, lib().sg_last_error(None).decode())
        self._h = h
        n = levels
        s = np.zeros(n, np.float32)
        w, hh, pt, b = (np.zeros(n, np.int32) for _ in range(4))
        self._check(lib().sg_get_geometry(self._h, s.ctypes, w.ctypes, hh.ctypes, pt.ctypes, b.ctypes))
        self.scales, self.widths, self.heights, self.pitches, self.budgets = s, w, hh, pt, b
        self.levels = levels
        self.cap = int(lib().sg_keypoint_capacity(self._h))

    def _check(self, rc):
        if rc != 0:
            raise SlamGpuError(rc, lib().sg_last_error(self._h).decode())

    def close(self):
        if self._h:
            # objects created from the context (databases, vocabularies) use its device, streams and pool: they go first
            for child in list(self.__dict__.get("_children", ())):
                child.close()
            lib().sg_destroy(self._h)
            self._h = None

    def _adopt(self, child):
        import weakref
        self.__dict__.setdefault("_children", weakref.WeakSet()).add(child)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing -----------------------------------------------------------------------------
    def synchronize(self):
        self._check(lib().sg_synchronize(self._h))

    def launch_count(self):
        return int(lib().sg_launch_count(self._h))

    def timer_start(self):
        self._check(lib().sg_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        self._check(lib().sg_timer_stop(self._h, C.byref(ms)))
        return ms.value

    STAGES = ("pyramid", "fast", "distribute", "describe", "match_topk", "match_resolve")

    def set_profiling(self, on=True):
        self._check(lib().sg_set_profiling(self._h, int(on)))

    def stage_ms(self):
        """Average per-call duration of every stage (ms) since set_profiling(True); -1 = did not run."""
        ms = (C.c_float * 6)()
        n = C.c_int()
        self._check(lib().sg_get_stage_ms(self._h, ms, C.byref(n)))
        return {k: float(v) for k, v in zip(self.STAGES, ms)}

    def flush_l2(self):
        self._check(lib().sg_flush_l2(self._h))

    def microbench_popc(self):
        r, ms = C.c_double(), C.c_float()
        self._check(lib().sg_microbench_popc(self._h, C.byref(r), C.byref(ms)))
        return r.value, ms.value

    def device_buffer(self, nbytes):
        return DeviceBuffer(self, nbytes)

    @staticmethod
    def _frames(imgs):
        imgs = np.asarray(imgs)
        if imgs.ndim == 2:
            imgs = imgs[None]
        assert imgs.dtype == np.uint8 and imgs.ndim == 3
        if imgs.strides[2] != 1 or imgs.strides[1] < imgs.shape[2] or imgs.strides[0] < imgs.strides[1] * imgs.shape[1]:
            imgs = np.ascontiguousarray(imgs)
        return imgs

    # ---- ImagePyramid (image_pyramid.hpp:16-30) ------------------------------------------------
    def pyramid_update(self, imgs):
        imgs = self._frames(imgs)
        assert imgs.shape[1] == self.params.height and imgs.shape[2] == self.params.width
        self._check(lib().sg_pyramid_update(self._h, imgs.ctypes.data, imgs.strides[1], imgs.strides[0], imgs.shape[0]))
        self._n = imgs.shape[0]

    def pyramid_update_device(self, dptr, pitch, frame_stride, n_frames):
        self._check(lib().sg_pyramid_update_device(self._h, dptr, pitch, frame_stride, n_frames))
        self._n = n_frames

    def get_level(self, frame, level, blurred=False):
        out = np.empty((int(self.heights[level]), int(self.widths[level])), np.uint8)
        self._check(lib().sg_pyramid_download(self._h, frame, level, int(blurred), out.ctypes.data, out.strides[0]))
        return out

    def get_blurred_level(self, frame, level):
        return self.get_level(frame, level, True)

    # ---- FeatureDetector (feature_detector.hpp:15-24) -------------------------------------------
    def detect(self):
        self._check(lib().sg_detect(self._h))

    def detected(self, frame, level, candidates=False):
        fn = lib().sg_detect_download_candidates if candidates else lib().sg_detect_download
        n = C.c_int()
        self._check(fn(self._h, frame, level, None, None, None, 0, C.byref(n)))
        m = n.value
        x, y, r = (np.empty(max(m, 1), np.int32) for _ in range(3))
        self._check(fn(self._h, frame, level, x.ctypes.data, y.ctypes.data, r.ctypes.data, m, C.byref(n)))
        return x[:m], y[:m], r[:m]

    # ---- OrbExtractor (orb_extractor.hpp:11-30) -------------------------------------------------
    def _alloc_out(self, n_frames, pinned=False):
        cap, lv = self.cap, self.levels
        spec = dict(x=((n_frames, cap), np.float32), y=((n_frames, cap), np.float32), angle=((n_frames, cap), np.float32),
                    octave=((n_frames, cap), np.int32), desc=((n_frames, cap, 8), np.uint32),
                    track_id=((n_frames, cap), np.int32), lvl_x=((n_frames, cap), np.int32),
                    lvl_y=((n_frames, cap), np.int32), count=((n_frames,), np.int32),
                    level_count=((n_frames, lv), np.int32))
        if pinned:   # page-locked host memory: the D2H copies of sg_extract run asynchronously
            pins = {k: PinnedArray(s, d) for k, (s, d) in spec.items()}
            arrs = {k: v.array for k, v in pins.items()}
            self.__dict__.setdefault("_pins", []).append(pins)   # keeps the allocations alive as long as the context
        else:
            arrs = {k: np.empty(s, d) for k, (s, d) in spec.items()}
        ks = Keypoints(*[arrs[k].ctypes.data for k, _ in Keypoints._fields_])
        return arrs, ks

    def set_overlap(self, parts):
        self._check(lib().sg_set_overlap(self._h, int(parts)))

    def set_pipeline_chunk(self, frames):
        self._check(lib().sg_set_pipeline_chunk(self._h, int(frames)))

    @staticmethod
    def _split(arrs, n_frames):
        out = []
        for f in range(n_frames):
            n = int(arrs["count"][f])
            d = {k: arrs[k][f, :n].copy() for k in ("x", "y", "angle", "octave", "desc", "track_id", "lvl_x", "lvl_y")}
            d["n"] = n
            d["level_counts"] = arrs["level_count"][f].copy()
            out.append(d)
        return out

    def detect_and_extract(self, imgs, tracks=None, track_ids=None):
        """OrbExtractor::detectAndExtract for a batch of frames.  tracks: optional list (one entry per
        frame) of (n, 2) float arrays of full-resolution tracker points."""
        imgs = self._frames(imgs)
        nf = imgs.shape[0]
        arrs, ks = self._alloc_out(nf)
        txy = tids = tn = None
        if tracks is not None and self.params.max_tracks > 0:
            T = self.params.max_tracks
            txy = np.zeros((nf, T, 2), np.float32)
            tids = np.zeros((nf, T), np.int32)
            tn = np.zeros(nf, np.int32)
            for f in range(nf):
                t = np.asarray(tracks[f], np.float32).reshape(-1, 2)
                tn[f] = len(t)
                txy[f, :len(t)] = t
                tids[f, :len(t)] = np.arange(len(t)) if track_ids is None else np.asarray(track_ids[f], np.int32)
        self._check(lib().sg_extract(self._h, imgs.ctypes.data, imgs.strides[1], imgs.strides[0], nf,
                                     None if txy is None else txy.ctypes.data, None if tids is None else tids.ctypes.data,
                                     None if tn is None else tn.ctypes.data, C.byref(ks)))
        self._n = nf
        return self._split(arrs, nf)

    def alloc_outputs(self, n_frames, pinned=True):
        """Host output arrays of a batch (dict of numpy arrays) and the sg_keypoints struct that points at them."""
        return self._alloc_out(n_frames, pinned)

    def extract_submit(self, imgs, base_frame, out_struct):
        """Streaming sg_extract: queue the batch on frame slots [base_frame, base_frame + len(imgs)) -> ticket."""
        t = C.c_int(-1)
        self._check(lib().sg_extract_submit(self._h, imgs.ctypes.data, imgs.strides[1], imgs.strides[0], imgs.shape[0],
                                            int(base_frame), C.byref(out_struct), C.byref(t)))
        return t.value

    def extract_wait(self, ticket):
        self._check(lib().sg_extract_wait(self._h, int(ticket)))

    def extract_device(self, dptr, pitch, frame_stride, n_frames):
        self._check(lib().sg_extract_device(self._h, dptr, pitch, frame_stride, n_frames))
        self._n = n_frames

    def extract_download(self, n_frames, only_counts=False):
        arrs, ks = self._alloc_out(n_frames)
        if only_counts:
            ks = Keypoints(None, None, None, None, None, None, None, None, arrs["count"].ctypes.data,
                           arrs["level_count"].ctypes.data)
        self._check(lib().sg_extract_download(self._h, n_frames, C.byref(ks)))
        if only_counts:
            return arrs["count"].copy(), arrs["level_count"].copy()
        return self._split(arrs, n_frames)

    def device_views(self):
        v = KeypointsDev()
        self._check(lib().sg_extract_device_views(self._h, C.byref(v)))
        return v

    # ---- matching -------------------------------------------------------------------------------
    def hamming(self, a, b):
        a = np.ascontiguousarray(a, np.uint32).reshape(-1, 8)
        b = np.ascontiguousarray(b, np.uint32).reshape(-1, 8)
        out = np.empty(len(a), np.uint32)
        self._check(lib().sg_hamming(self._h, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data))
        return out

    def match_bruteforce(self, dA, aA, dB, aB, ratio=0.8, thr=50, check_orientation=True, ratio_is_double=False):
        """Brute-force case of matchForLoopClosures: returns (num_matches, matches[nA])."""
        dA = np.ascontiguousarray(dA, np.uint32).reshape(-1, 8)
        dB = np.ascontiguousarray(dB, np.uint32).reshape(-1, 8)
        aA = np.ascontiguousarray(aA, np.float32)
        aB = np.ascontiguousarray(aB, np.float32)
        mp = MatchParams(ratio, thr, int(check_orientation), int(ratio_is_double))
        m = np.empty(max(len(dA), 1), np.int32)
        n = C.c_uint32()
        self._check(lib().sg_match_bruteforce(self._h, dA.ctypes.data, aA.ctypes.data, len(dA), dB.ctypes.data,
                                              aB.ctypes.data, len(dB), C.byref(mp), m.ctypes.data, C.byref(n)))
        return int(n.value), m[:len(dA)]

    def match_bow(self, dA, aA, nodeA, dB, aB, nodeB, eligA=None, eligB=None, ratio=0.8, thr=50, check_orientation=True,
                  ratio_is_double=False):
        """matchForLoopClosures with DBoW2 node buckets: returns (num_matches, matches[nA])."""
        dA = np.ascontiguousarray(dA, np.uint32).reshape(-1, 8); dB = np.ascontiguousarray(dB, np.uint32).reshape(-1, 8)
        aA = np.ascontiguousarray(aA, np.float32); aB = np.ascontiguousarray(aB, np.float32)
        nodeA = np.ascontiguousarray(nodeA, np.int32); nodeB = np.ascontiguousarray(nodeB, np.int32)
        eA = None if eligA is None else np.ascontiguousarray(eligA, np.uint8)
        eB = None if eligB is None else np.ascontiguousarray(eligB, np.uint8)
        mp = MatchParams(ratio, thr, int(check_orientation), int(ratio_is_double))
        m = np.empty(max(len(dA), 1), np.int32)
        n = C.c_uint32()
        p = lambda a: None if a is None else a.ctypes.data
        self._check(lib().sg_match_bow(self._h, p(dA), p(aA), p(nodeA), p(eA), len(dA), p(dB), p(aB), p(nodeB), p(eB), len(dB),
                                       C.byref(mp), m.ctypes.data, C.byref(n)))
        return int(n.value), m[:len(dA)]

    def match_triangulation(self, dA, aA, octA, bearA, nodeA, dB, aB, bearB, nodeB, E, scale_factors, eligA=None, eligB=None,
                            residual_deg_thr=0.2, thr=50, check_orientation=True):
        """matchForTriangulationDBoW: returns (num_matches, matches[nA])."""
        dA = np.ascontiguousarray(dA, np.uint32).reshape(-1, 8); dB = np.ascontiguousarray(dB, np.uint32).reshape(-1, 8)
        aA = np.ascontiguousarray(aA, np.float32); aB = np.ascontiguousarray(aB, np.float32)
        octA = np.ascontiguousarray(octA, np.int32)
        bearA = np.ascontiguousarray(bearA, np.float64); bearB = np.ascontiguousarray(bearB, np.float64)
        nodeA = np.ascontiguousarray(nodeA, np.int32); nodeB = np.ascontiguousarray(nodeB, np.int32)
        eA = None if eligA is None else np.ascontiguousarray(eligA, np.uint8)
        eB = None if eligB is None else np.ascontiguousarray(eligB, np.uint8)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        tp = TriangulationParams()
        for i, x in enumerate(np.asarray(E, np.float64).reshape(9)):
            tp.E[i] = float(x)
        tp.scale_factors = sf.ctypes.data; tp.n_levels = len(sf); tp.residual_deg_thr = residual_deg_thr
        tp.thr = thr; tp.check_orientation = int(check_orientation)
        m = np.empty(max(len(dA), 1), np.int32)
        n = C.c_uint32()
        p = lambda a: None if a is None else a.ctypes.data
        self._check(lib().sg_match_triangulation(self._h, p(dA), p(aA), p(octA), p(bearA), p(nodeA), p(eA), len(dA), p(dB), p(aB),
                                                 p(bearB), p(nodeB), p(eB), len(dB), C.byref(tp), m.ctypes.data, C.byref(n)))
        return int(n.value), m[:len(dA)]

    def rescans(self):
        return int(lib().sg_match_rescans(self._h))

    # ---- candidate-list matchers and descriptor medoid ------------------------------------
    def search_candidates(self, kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=0, thr=50, taken=None, qlevel=None, order=None):
        """Radius query + best (/ second best) Hamming per query; returns (n_matched, idx[nQ], dist[nQ]);
        `taken` (uint8, mode 1) is updated in place."""
        kx = np.ascontiguousarray(kx, np.float32); ky = np.ascontiguousarray(ky, np.float32)
        koct = np.ascontiguousarray(koct, np.int32); kdesc = np.ascontiguousarray(kdesc, np.uint32).reshape(-1, 8)
        qx = np.ascontiguousarray(qx, np.float32); qy = np.ascontiguousarray(qy, np.float32)
        qr = np.ascontiguousarray(qr, np.float32); qdesc = np.ascontiguousarray(qdesc, np.uint32).reshape(-1, 8)
        ql = None if qlevel is None else np.ascontiguousarray(qlevel, np.int32)
        od = None if order is None else np.ascontiguousarray(order, np.int32)
        if taken is not None:
            assert taken.dtype == np.uint8 and taken.flags.c_contiguous and len(taken) == len(kx)
        idx = np.empty(max(len(qx), 1), np.int32)
        dist = np.empty(max(len(qx), 1), np.uint32)
        n = C.c_uint32()
        self._check(lib().sg_search_candidates(
            self._h, kx.ctypes.data, ky.ctypes.data, koct.ctypes.data, kdesc.ctypes.data, len(kx),
            None if od is None else od.ctypes.data, None if taken is None else taken.ctypes.data,
            qx.ctypes.data, qy.ctypes.data, qr.ctypes.data, qdesc.ctypes.data, None if ql is None else ql.ctypes.data,
            len(qx), int(mode), int(thr), idx.ctypes.data, dist.ctypes.data, C.byref(n)))
        return int(n.value), idx[:len(qx)], dist[:len(qx)]

    def match_sim3(self, x1, y1, oct1, d1, x2, y2, oct2, d2, q12, q12desc, q12lvl, q21, q21desc, q21lvl, order1=None, order2=None):
        """matchMapPointsSim3 on projected queries (q12 [n1, 3] = x, y, r of keyframe 1's map points in keyframe 2, r < 0: no
        query; q21 [n2, 3] the reverse); returns the agreed (i, j) keypoint pairs [n, 2]."""
        f = lambda a: np.ascontiguousarray(a, np.float32)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        u = lambda a: np.ascontiguousarray(a, np.uint32).reshape(-1, 8)
        x1, y1, x2, y2 = f(x1), f(y1), f(x2), f(y2)
        oct1, oct2, q12lvl, q21lvl = i32(oct1), i32(oct2), i32(q12lvl), i32(q21lvl)
        d1, d2, q12desc, q21desc = u(d1), u(d2), u(q12desc), u(q21desc)
        q12 = f(q12).reshape(-1, 3); q21 = f(q21).reshape(-1, 3)
        a = [np.ascontiguousarray(q12[:, k]) for k in range(3)] + [np.ascontiguousarray(q21[:, k]) for k in range(3)]
        o1 = None if order1 is None else i32(order1)
        o2 = None if order2 is None else i32(order2)
        pairs = np.zeros((max(min(len(x1), len(x2)), 1), 2), np.int32)
        n = C.c_uint32()
        p = lambda v: None if v is None else v.ctypes.data
        self._check(lib().sg_match_sim3(self._h, p(x1), p(y1), p(oct1), p(d1), len(x1), p(o1), p(x2), p(y2), p(oct2), p(d2), len(x2),
                                        p(o2), p(a[0]), p(a[1]), p(a[2]), p(q12desc), p(q12lvl), p(a[3]), p(a[4]), p(a[5]),
                                        p(q21desc), p(q21lvl), p(pairs), C.byref(n)))
        return pairs[:n.value].copy()

    def medoid(self, desc, offsets):
        desc = np.ascontiguousarray(desc, np.uint32).reshape(-1, 8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        best = np.empty(max(len(offsets) - 1, 1), np.int32)
        self._check(lib().sg_medoid(self._h, desc.ctypes.data, offsets.ctypes.data, len(offsets) - 1, best.ctypes.data))
        return best[:len(offsets) - 1]


class DescriptorDB:
    """Device-resident descriptor sets (one per keyframe) for batched pair matching."""

    def __init__(self, ctx, desc, angle, offsets=None, device_ptrs=None, view=False):
        """device_ptrs=(d_desc, d_angle): build from device arrays; view=True reads them in place instead of copying."""
        self.ctx = ctx
        h = C.c_void_p()
        if device_ptrs is not None:
            d_desc, d_angle = device_ptrs
            offsets = np.ascontiguousarray(offsets, np.int64)
            fn = lib().sg_db_wrap_device if view else lib().sg_db_create_device
            ctx._check(fn(ctx._h, d_desc, d_angle, offsets.ctypes.data, len(offsets) - 1, C.byref(h)))
        else:
            desc = np.ascontiguousarray(desc, np.uint32)
            angle = np.ascontiguousarray(angle, np.float32)
            if offsets is None:   # (n_sets, n_per_set, 8)
                n_sets, per = desc.shape[0], desc.shape[1]
                offsets = np.arange(n_sets + 1, dtype=np.int64) * per
            offsets = np.ascontiguousarray(offsets, np.int64)
            ctx._check(lib().sg_db_create(ctx._h, desc.ctypes.data, angle.ctypes.data, offsets.ctypes.data,
                                          len(offsets) - 1, C.byref(h)))
        self._h = h
        ctx._adopt(self)
        self.offsets = offsets
        self.max_set = int(np.max(np.diff(offsets))) if len(offsets) > 1 else 0

    def match_pairs(self, pairs, ratio=0.8, thr=50, check_orientation=True, ratio_is_double=False, want_matches=True, out=None):
        """out: optional preallocated (counts uint32[n_pairs], rows int32[n_pairs, max_set]) -- e.g. PinnedArray.array."""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        mp = MatchParams(ratio, thr, int(check_orientation), int(ratio_is_double))
        if out is not None:
            n, m = out
            assert n.dtype == np.uint32 and len(n) >= len(pairs) and m.dtype == np.int32 and m.shape[0] >= len(pairs)
            assert m.shape[1] == max(self.max_set, 1) and m.flags.c_contiguous
        else:
            n = np.zeros(len(pairs), np.uint32)
            m = np.empty((len(pairs), max(self.max_set, 1)), np.int32) if want_matches else None
        self.ctx._check(lib().sg_match_pairs(self.ctx._h, self._h, pairs.ctypes.data, len(pairs), C.byref(mp),
                                             None if m is None else m.ctypes.data, max(self.max_set, 1), n.ctypes.data))
        return n, m

    def match_pairs_device(self, d_pairs, n_pairs, d_counts, d_matches=None, ratio=0.8, thr=50, check_orientation=True):
        mp = MatchParams(ratio, thr, int(check_orientation), 0)
        self.ctx._check(lib().sg_match_pairs_device(self.ctx._h, self._h, d_pairs, n_pairs, C.byref(mp), d_matches,
                                                    max(self.max_set, 1), d_counts))

    def close(self):
        if self._h:
            lib().sg_db_destroy(self._h)
            self._h = None


class Vocabulary:
    """Device-resident DBoW2-shaped vocabulary tree (see synth.random_vocabulary for the array layout)."""

    def __init__(self, ctx, vocab):
        self.ctx = ctx
        a = {k: np.ascontiguousarray(vocab[k], t) for k, t in (("child_off", np.int32), ("child_ids", np.int32),
             ("node_desc", np.uint32), ("node_weight", np.float64), ("node_word", np.int32))}
        h = C.c_void_p()
        ctx._check(lib().sg_vocab_create(ctx._h, a["child_off"].ctypes.data, a["child_ids"].ctypes.data, a["node_desc"].ctypes.data,
                                         a["node_weight"].ctypes.data, a["node_word"].ctypes.data, len(a["node_word"]),
                                         int(vocab["levels"]), C.byref(h)))
        self._h = h
        ctx._adopt(self)

    def transform(self, desc, levels_up=4):
        desc = np.ascontiguousarray(desc, np.uint32).reshape(-1, 8)
        n = len(desc)
        word = np.empty(max(n, 1), np.int32); weight = np.empty(max(n, 1), np.float64); node = np.empty(max(n, 1), np.int32)
        self.ctx._check(lib().sg_bow_transform(self.ctx._h, self._h, desc.ctypes.data, n, int(levels_up), word.ctypes.data,
                                               weight.ctypes.data, node.ctypes.data))
        return word[:n], weight[:n], node[:n]

    def bow_vector(self, word, weight):
        """BowVector (ascending words, L1-normalised double values) from the per-feature output of transform()."""
        word = np.ascontiguousarray(word, np.int32); weight = np.ascontiguousarray(weight, np.float64)
        n = len(word)
        vw = np.empty(max(n, 1), np.uint32); vv = np.empty(max(n, 1), np.float64); k = C.c_int(0)
        self.ctx._check(lib().sg_bow_vector(self.ctx._h, word.ctypes.data, weight.ctypes.data, n, vw.ctypes.data, vv.ctypes.data,
                                            C.byref(k)))
        return vw[:k.value].copy(), vv[:k.value].copy()

    def bow_vector_batch(self, word, weight, offsets):
        """BowVectors of many keyframes in one launch: keyframe k owns features offsets[k]..offsets[k+1].
        -> list of (words, values) per keyframe."""
        word = np.ascontiguousarray(word, np.int32); weight = np.ascontiguousarray(weight, np.float64)
        offsets = np.ascontiguousarray(offsets, np.int64)
        nk = len(offsets) - 1
        vw = np.empty(max(len(word), 1), np.uint32); vv = np.empty(max(len(word), 1), np.float64); cnt = np.zeros(max(nk, 1), np.int32)
        self.ctx._check(lib().sg_bow_vector_batch(self.ctx._h, word.ctypes.data, weight.ctypes.data, offsets.ctypes.data, nk,
                                                  vw.ctypes.data, vv.ctypes.data, cnt.ctypes.data))
        return [(vw[offsets[k]:offsets[k] + cnt[k]].copy(), vv[offsets[k]:offsets[k] + cnt[k]].copy()) for k in range(nk)]

    def close(self):
        if self._h:
            lib().sg_vocab_destroy(self._h)
            self._h = None


class BowDatabase:
    """Device-resident BowVectors of the keyframes (BowIndex::add / remove / getBowSimilar, bow_index.cpp:44-176)."""

    def __init__(self, ctx, max_keyframes, max_words_per_keyframe=2048):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._check(lib().sg_bowdb_create(ctx._h, int(max_keyframes), int(max_words_per_keyframe), C.byref(h)))
        self._h = h
        ctx._adopt(self)

    def __len__(self):
        return int(lib().sg_bowdb_size(self._h))

    def add(self, map_id, kf_id, vec_word, vec_value):
        vw = np.ascontiguousarray(vec_word, np.uint32); vv = np.ascontiguousarray(vec_value, np.float64)
        self.ctx._check(lib().sg_bowdb_add(self.ctx._h, self._h, int(map_id), int(kf_id), vw.ctypes.data, vv.ctypes.data, len(vw)))

    def remove(self, map_id, kf_id):
        self.ctx._check(lib().sg_bowdb_remove(self.ctx._h, self._h, int(map_id), int(kf_id)))

    def similar(self, vec_word, vec_value, self_key=(-1, -1), min_in_common_ratio=0.8, score_ratio=0.75, capacity=None):
        """-> (map ids, keyframe ids, scores), best first."""
        vw = np.ascontiguousarray(vec_word, np.uint32); vv = np.ascontiguousarray(vec_value, np.float64)
        cap = max(len(self), 1) if capacity is None else int(capacity)
        om = np.empty(cap, np.int32); ok = np.empty(cap, np.int32); osc = np.empty(cap, np.float32); n = C.c_int(0)
        self.ctx._check(lib().sg_bow_similar(self.ctx._h, self._h, vw.ctypes.data, vv.ctypes.data, len(vw), int(self_key[0]),
                                             int(self_key[1]), C.c_float(min_in_common_ratio), C.c_float(score_ratio),
                                             om.ctypes.data, ok.ctypes.data, osc.ctypes.data, cap, C.byref(n)))
        k = min(n.value, cap)
        return om[:k].copy(), ok[:k].copy(), osc[:k].copy()

    def close(self):
        if self._h:
            lib().sg_bowdb_destroy(self._h)
            self._h = None


# ---- mirrors of the reference's class names (thin; one Context underneath) ------------------------
class OrbExtractor:
    """OrbExtractor::build(settings) + detectAndExtract (orb_extractor.hpp:16-22)."""

    def __init__(self, ctx):
        self.ctx = ctx

    @staticmethod
    def build(width, height, **kw):
        return OrbExtractor(Context(width, height, **kw))

    def detect_and_extract(self, imgs, tracks=None, track_ids=None):
        return self.ctx.detect_and_extract(imgs, tracks, track_ids)


def match_for_loop_closures(ctx, kf1, kf2, ratio=0.8, check_orientation=True):
    """matchForLoopClosures (keyframe_matcher.hpp:33-40) on two extraction results (dicts with `desc` and
    `angle`), all features in one BoW node and owning triangulated map points."""
    return ctx.match_bruteforce(kf1["desc"], kf1["angle"], kf2["desc"], kf2["angle"], ratio=ratio, thr=50,
                                check_orientation=check_orientation)
