"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product path
(scalable-video-codec_b200/) never does.

Two checkers live here:
  * ``liboracle.so``   -- oracle/svc_oracle.c, the plain-C restatement.
  * ``_ref/libref_motion.so`` -- the UNMODIFIED reference libs/motion.cpp
    (compiled by oracle/Makefile where /root/reference exists; the built .so
    travels to the GPU box, the sources never enter this repository).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (and _ref when the reference sources are present)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(p):
            build()
        _lib = C.CDLL(p)
        _lib.orc_padded_dim.restype = C.c_uint
        _lib.orc_serialized_frame_bytes.restype = C.c_uint64
        _lib.orc_dct_planar.restype = C.c_int
    return _lib


def ref():
    """The compiled unmodified reference, or None if it was never built."""
    global _ref
    if _ref is None:
        _ref = _load(os.path.join(_HERE, "_ref", "libref_motion.so"))
    return _ref


def have_ref() -> bool:
    return ref() is not None


def _ptr(a, ty=_u8p):
    return a.ctypes.data_as(ty)


def _ptr_array(arrs):
    arr = (_u8p * len(arrs))()
    for i, a in enumerate(arrs):
        assert a.dtype == np.uint8 and a.flags.c_contiguous
        arr[i] = _ptr(a)
    return arr


def padded_dim(a, block, levels):
    return int(lib().orc_padded_dim(C.c_uint(a), C.c_uint(block),
                                    C.c_uint(1 << (levels - 1))))


def bgr_to_y(bgr, pw, ph):
    h, w, _ = bgr.shape
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    out = np.empty((ph, pw), np.uint8)
    lib().orc_bgr_to_y(_ptr(bgr), w, h, pw, ph, _ptr(out))
    return out


def pyr_down(src):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    h, w = src.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyr_down(_ptr(src), w, h, _ptr(out))
    return out


def y_pyramid(bgr, pw, ph, levels):
    """[level0 .. level(levels-1)] tightly packed uint8 planes."""
    h, w, _ = bgr.shape
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    outs = [np.empty((ph >> l, pw >> l), np.uint8) for l in range(levels)]
    lib().orc_y_pyramid(_ptr(bgr), w, h, pw, ph, levels, _ptr_array(outs))
    return outs


def _mv_out(fw, fh, bw, bh):
    mw, mh = fw // bw, fh // bh
    return np.empty((mh, mw, 2), np.float32), np.empty((mh, mw), np.float32)


def ebma(tracked, anchor, r, bw, bh, impl="oracle"):
    fh, fw = tracked.shape
    mv, mad = _mv_out(fw, fh, bw, bh)
    fn = lib().orc_ebma if impl == "oracle" else ref().ref_ebma
    fn(_ptr(np.ascontiguousarray(tracked)), _ptr(np.ascontiguousarray(anchor)),
       fw, fh, r, bw, bh, _ptr(mv, _f32p), _ptr(mad, _f32p))
    return mv, mad


def hbma(tracked_pyr, anchor_pyr, search_range, bw=16, bh=16, impl="oracle"):
    """impl: 'oracle' (C restatement), 'ref' (reference generic),
    'ref_sse2' (reference EstimateMotionHierarchical16x16Sse2)."""
    levels = len(tracked_pyr)
    fh, fw = tracked_pyr[0].shape
    mv, mad = _mv_out(fw, fh, bw, bh)
    t, a = _ptr_array(tracked_pyr), _ptr_array(anchor_pyr)
    if impl == "oracle":
        lib().orc_hbma(t, a, levels, fw, fh, search_range, bw, bh,
                       _ptr(mv, _f32p), _ptr(mad, _f32p))
    elif impl == "ref":
        ref().ref_hbma(t, a, levels, fw, fh, search_range, bw, bh,
                       _ptr(mv, _f32p), _ptr(mad, _f32p))
    elif impl == "ref_sse2":
        assert levels == 4 and bw == 16 and bh == 16
        ref().ref_hbma_16x16_sse2(t, a, fw, fh, search_range,
                                  _ptr(mv, _f32p), _ptr(mad, _f32p))
    else:
        raise ValueError(impl)
    return mv, mad


def hbma_count(tracked_pyr, anchor_pyr, search_range, bw=16, bh=16):
    levels = len(tracked_pyr)
    fh, fw = tracked_pyr[0].shape
    nc, na = C.c_uint64(), C.c_uint64()
    lib().orc_hbma_count(_ptr_array(tracked_pyr), _ptr_array(anchor_pyr),
                         levels, fw, fh, search_range, bw, bh,
                         C.byref(nc), C.byref(na))
    return nc.value, na.value


def dct_planar(bgr, pw, ph, tbw=8, tbh=8):
    """(3, ph, pw) float32: B, G, R coefficient planes of the zero-padded frame."""
    h, w, _ = bgr.shape
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    out = np.empty((3, ph, pw), np.float32)
    planes = (_f32p * 3)(*[_ptr(out[c], _f32p) for c in range(3)])
    rc = lib().orc_dct_planar(_ptr(bgr), w, h, pw, ph, tbw, tbh, planes)
    if rc != 0:
        raise ValueError("orc_dct_planar: bad block size")
    return out


def header(n_input_frames, w, h, pw, ph, tbw=8, tbh=8, channels=3):
    out = np.empty(32, np.uint8)
    lib().orc_header(n_input_frames, w, h, pw - w, ph - h, tbw, tbh, channels,
                     _ptr(out))
    return out


def serialized_frame_bytes(w, h, tbw=8, tbh=8, channels=3):
    return int(lib().orc_serialized_frame_bytes(w, h, tbw, tbh, channels))


def serialize_frame(planes, block_types, w, h, tbw, tbh, mv_field_w,
                    mv_block_w, mv_block_h):
    """planes: (C, ph, pw) float32 contiguous. block_types: uint32 or None."""
    planes = np.ascontiguousarray(planes, dtype=np.float32)
    ch = planes.shape[0]
    n = serialized_frame_bytes(w, h, tbw, tbh, ch)
    out = np.empty(n, np.uint8)
    pp = (_f32p * ch)(*[_ptr(planes[c], _f32p) for c in range(ch)])
    bt = None
    if block_types is not None:
        block_types = np.ascontiguousarray(block_types, dtype=np.uint32)
        bt = block_types.ctypes.data_as(C.POINTER(C.c_uint32))
    lib().orc_serialize_frame(pp, C.c_uint64(planes[0].size), ch, bt, w, h,
                              tbw, tbh, mv_field_w, mv_block_w, mv_block_h,
                              _ptr(out))
    return out


def gaze_rect(gx, gy, max_w, max_h, fw, fh, pw, ph):
    out = (C.c_uint * 4)()
    lib().orc_gaze_rect(gx, gy, max_w, max_h, fw, fh, pw, ph, out)
    return tuple(int(v) for v in out)


def decode_frame_blocks(records, pw, ph, tbw=8, tbh=8, fg_q=1, bg_q=640, gaze=None):
    """(ph, pw, 3) float32: the decoder's `upscaled_frame` before /255 and resize."""
    records = np.ascontiguousarray(records, dtype=np.uint8)
    out = np.empty((ph, pw, 3), np.float32)
    g = (C.c_uint * 4)(*gaze) if gaze is not None else None
    lib().orc_decode_frame_blocks.restype = C.c_int
    rc = lib().orc_decode_frame_blocks(_ptr(records), pw, ph, tbw, tbh, fg_q, bg_q, g, _ptr(out, _f32p))
    if rc != 0:
        raise ValueError("orc_decode_frame_blocks: bad block size")
    return out


# ---- block-type stages (libs/encoder.cpp:491-624): the checkers are the compiled reference
# ---- (RANSAC) and python cv2 (morphologyEx, kmeans, connectedComponents) -------------------
_ref_copies = {}


def ref_seeded(seed: int, fresh: bool = True):
    """A private copy of the compiled reference whose function-static RANSAC engine
    (libs/motion.cpp:186-187) will be seeded with `seed` on its first call: the engine is
    seeded once per loaded library image, so every request gets its own image (fresh=False
    returns the image already loaded for that seed, with its engine state carried on)."""
    if not fresh and seed in _ref_copies:
        return _ref_copies[seed]
    import shutil
    import tempfile
    src = os.path.join(_HERE, "_ref", "libref_motion.so")
    if not os.path.exists(src):
        return None
    d = tempfile.mkdtemp(prefix="svc_ref_")
    p = os.path.join(d, "libref_motion_%d_%d.so" % (seed, len(_ref_copies)))
    shutil.copy(src, p)
    L = C.CDLL(p)
    if not hasattr(L, "ref_ransac"):
        return None
    L.ref_set_seed(C.c_uint(seed))
    _ref_copies[seed] = L
    _ref_copies[("all", len(_ref_copies))] = L  # keep every image alive
    return L


def ref_ransac(L, mv, subset_sz=1, inlier_thresh=7.5, success_prob=0.99, inlier_ratio=0.5, gm0=(0.0, 0.0)):
    """EstimateGlobalMotionRansac of the compiled reference image `L` (see ref_seeded)."""
    mv = np.ascontiguousarray(mv, np.float32).reshape(-1, 2)
    n = mv.shape[0]
    buf = np.zeros((n + 1, 2), np.float32)  # the reference may read index n (libs/motion.cpp:208)
    buf[:n] = mv
    buf[n] = np.nan
    rm, ni = C.c_float(), C.c_uint32()
    gm = np.array(gm0, np.float32)
    inl = np.zeros(n, np.uint32)
    L.ref_ransac(buf.ctypes.data_as(_f32p), n, subset_sz, C.c_float(inlier_thresh), C.c_float(success_prob),
                 C.c_float(inlier_ratio), C.byref(rm), gm.ctypes.data_as(_f32p),
                 inl.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(ni))
    return rm.value, gm, inl[:ni.value].copy()


def block_types_cv2(mv_field, inliers, kmeans_seed, morph_rect=(3, 3), cluster_count=10, attempts=3,
                    max_iter=10, eps=1.0, connectivity=4, mv_block=(16, 16)):
    """libs/encoder.cpp:507-624 with python cv2 standing in for the C++ OpenCV calls, given the
    RANSAC inliers.  Returns block types (h x w, uint32)."""
    import cv2
    mv = np.asarray(mv_field, np.float32)
    h, w = mv.shape[:2]
    mask = np.full((h, w), 255, np.uint8)
    mask.reshape(-1)[np.asarray(inliers, np.int64)] = 0
    el = cv2.getStructuringElement(cv2.MORPH_RECT, morph_rect)
    mask = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, el)
    mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, el)
    fg = np.flatnonzero(mask.reshape(-1) == 255)
    types = np.zeros(h * w, np.uint32)
    if fg.size == 0:
        return types.reshape(h, w)
    k = min(cluster_count, fg.size)
    feats = np.zeros((fg.size, 1, 4), np.float32)  # BuildMvFeatures quirk: (0, mv.x, x, y)
    feats[:, 0, 1] = mv.reshape(-1, 2)[fg, 0]
    feats[:, 0, 2] = (fg % w) * mv_block[0]
    feats[:, 0, 3] = (fg // w) * mv_block[1]
    cv2.setRNGSeed(kmeans_seed)
    _, ids, _ = cv2.kmeans(feats, k, None, (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, max_iter, eps),
                           attempts, cv2.KMEANS_PP_CENTERS)
    ids = ids.reshape(-1)
    offset = 0
    for cid in range(k):
        cm = np.zeros(h * w, np.uint8)
        cm[fg[ids == cid]] = 255
        n, lab = cv2.connectedComponents(cm.reshape(h, w), connectivity=connectivity, ltype=cv2.CV_32S)
        lab = lab.reshape(-1)
        sel = fg[lab[fg] != 0]
        types[sel] = lab[sel] + offset
        offset += n
    return types.reshape(h, w)
