"""ctypes binding of include/svc_segment.h (exported by lib/libsvc_host.so): the block-type
stages that consume the motion field (libs/encoder.cpp:491-624).  Marshalling only."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)

MORPH_ERODE, MORPH_DILATE, MORPH_OPEN, MORPH_CLOSE = 0, 1, 2, 3


class SegError(ValueError):
    pass


class _Cfg(C.Structure):
    _fields_ = [("ransac_subset_sz", C.c_uint32), ("ransac_inlier_thresh", C.c_float),
                ("ransac_success_prob", C.c_float), ("ransac_inlier_ratio", C.c_float),
                ("morph_rect_w", C.c_uint32), ("morph_rect_h", C.c_uint32),
                ("kmeans_cluster_count", C.c_uint32), ("kmeans_attempt_count", C.c_uint32),
                ("kmeans_max_iter_count", C.c_uint32), ("kmeans_epsilon", C.c_float),
                ("connected_components_connectivity", C.c_uint32),
                ("mv_block_w", C.c_uint32), ("mv_block_h", C.c_uint32)]


@dataclass
class SegmentConfig:  # defaults: apps/encoder.cpp:28-58
    ransac_subset_sz: int = 1
    ransac_inlier_thresh: float = 7.5
    ransac_success_prob: float = 0.99
    ransac_inlier_ratio: float = 0.5
    morph_rect_w: int = 3
    morph_rect_h: int = 3
    kmeans_cluster_count: int = 10
    kmeans_attempt_count: int = 3
    kmeans_max_iter_count: int = 10
    kmeans_epsilon: float = 1.0
    connected_components_connectivity: int = 4
    mv_block_w: int = 16
    mv_block_h: int = 16

    def _c(self):
        return _Cfg(*[getattr(self, f) for f, _ in _Cfg._fields_])


def host_lib_path() -> str:
    return os.environ.get("SVC_B200_HOST_LIB", os.path.join(os.path.dirname(_PKG), "lib", "libsvc_host.so"))


_lib = None


def host_lib() -> C.CDLL:
    global _lib
    if _lib is None:
        p = host_lib_path()
        if not os.path.exists(p):
            raise ImportError(f"{p} not found: build it with `make -C scalable-video-codec_b200`")
        _lib = C.CDLL(p)
        _lib.svc_seg_last_error.restype = C.c_char_p
    return _lib


def _check(rc):
    if rc != 0:
        raise SegError(host_lib().svc_seg_last_error().decode())


def validate(cfg: SegmentConfig) -> str:
    """'' when valid, else the reference's Validate message (libs/encoder.cpp:20-101)."""
    L = host_lib()
    c = cfg._c()
    return "" if L.svc_seg_validate(C.byref(c)) == 0 else L.svc_seg_last_error().decode()


def ransac(mv, subset_sz=1, inlier_thresh=7.5, success_prob=0.99, inlier_ratio=0.5, rng_state=1, gm0=(0.0, 0.0)):
    """EstimateGlobalMotionRansac (libs/motion.hpp:99-103) -> (rmse, global_motion, inliers, new rng state)."""
    mv = np.ascontiguousarray(mv, np.float32).reshape(-1, 2)
    n = mv.shape[0]
    st, rm, ni = C.c_uint32(rng_state), C.c_float(), C.c_uint32()
    gm = np.array(gm0, np.float32)
    inl = np.zeros(max(n, 1), np.uint32)
    _check(host_lib().svc_seg_ransac(mv.ctypes.data_as(_f32p), n, subset_sz, C.c_float(inlier_thresh),
                                     C.c_float(success_prob), C.c_float(inlier_ratio), C.byref(st), C.byref(rm),
                                     gm.ctypes.data_as(_f32p), inl.ctypes.data_as(_u32p), C.byref(ni)))
    return rm.value, gm, inl[:ni.value].copy(), st.value


def global_motion_avg(mv):
    mv = np.ascontiguousarray(mv, np.float32).reshape(-1, 2)
    gm = np.zeros(2, np.float32)
    _check(host_lib().svc_seg_global_motion_avg(mv.ctypes.data_as(_f32p), mv.shape[0], gm.ctypes.data_as(_f32p)))
    return gm


def morphology(mask, op, rect_w=3, rect_h=3):
    m = np.ascontiguousarray(mask, np.uint8).copy()
    h, w = m.shape
    _check(host_lib().svc_seg_morphology(m.ctypes.data_as(_u8p), w, h, op, rect_w, rect_h))
    return m


def connected_components(mask, connectivity=4):
    m = np.ascontiguousarray(mask, np.uint8)
    h, w = m.shape
    lab = np.zeros((h, w), np.int32)
    n = C.c_uint32()
    _check(host_lib().svc_seg_connected_components(m.ctypes.data_as(_u8p), w, h, connectivity,
                                                   lab.ctypes.data_as(_i32p), C.byref(n)))
    return n.value, lab


def kmeans(data, k, max_iter=10, eps=1.0, attempts=3, rng_state=0xffffffff):
    """cv::kmeans(KMEANS_PP_CENTERS) -> (compactness, labels, centers, new cv::RNG state)."""
    d = np.ascontiguousarray(data, np.float32)
    n, dims = d.shape
    lab = np.zeros(n, np.int32)
    cen = np.zeros((k, dims), np.float32)
    comp, st = C.c_double(), C.c_uint64(rng_state)
    _check(host_lib().svc_seg_kmeans(d.ctypes.data_as(_f32p), n, dims, k, max_iter, C.c_float(eps), attempts,
                                     C.byref(st), lab.ctypes.data_as(_i32p), cen.ctypes.data_as(_f32p), C.byref(comp)))
    return comp.value, lab, cen, st.value


def block_types(mv_field, cfg: SegmentConfig = None, ransac_rng_state=1, kmeans_rng_state=0xffffffff):
    """libs/encoder.cpp:491-624 for one motion field (mv_field: h x w x 2) ->
    (block types h x w u32, global motion, new ransac state, new kmeans state)."""
    cfg = cfg or SegmentConfig()
    mv = np.ascontiguousarray(mv_field, np.float32)
    h, w = mv.shape[:2]
    bt = np.zeros((h, w), np.uint32)
    gm = np.zeros(2, np.float32)
    rs, ks = C.c_uint32(ransac_rng_state), C.c_uint64(kmeans_rng_state)
    c = cfg._c()
    _check(host_lib().svc_seg_block_types(mv.ctypes.data_as(_f32p), w, h, C.byref(c), C.byref(rs), C.byref(ks),
                                          bt.ctypes.data_as(_u32p), gm.ctypes.data_as(_f32p)))
    return bt, gm, rs.value, ks.value


def frame_generators(seed: int, frame: int):
    """(ransac rng state, kmeans rng state) svc::Encoder uses for encoded frame `frame`."""
    r, k = C.c_uint32(), C.c_uint64()
    host_lib().svc_seg_frame_generators(C.c_uint64(seed), C.c_uint64(frame), C.byref(r), C.byref(k))
    return r.value, k.value


def block_types_batch(mv_fields, cfg: SegmentConfig = None, seed=1, first_frame=0, threads=0):
    """svc::BlockTypeStage over n motion fields (n x h x w x 2) -> n x h x w block types."""
    cfg = cfg or SegmentConfig()
    mv = np.ascontiguousarray(mv_fields, np.float32)
    n, h, w = mv.shape[:3]
    bt = np.zeros((n, h, w), np.uint32)
    c = cfg._c()
    _check(host_lib().svc_seg_block_types_batch(mv.ctypes.data_as(_f32p), n, w, h, C.byref(c), C.c_uint64(seed),
                                                C.c_uint64(first_frame), threads, bt.ctypes.data_as(_u32p)))
    return bt
