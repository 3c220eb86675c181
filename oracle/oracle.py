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
