"""ctypes binding of include/svc_b200.h (one python function per C entry point).

Function names follow the reference interface they stand in for
(libs/motion.hpp:106-152, libs/encoder.cpp:222-269, 323-339, 459-470).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)

STAGE_Y_PYRAMID, STAGE_HBMA, STAGE_DCT_STREAM, STAGE_PYR_DOWN = 1, 2, 3, 4
# svc_session_config.hbma_kernel_family (test hook): SVC_HBMA_FAMILY_*
HBMA_FAMILY_AUTO, HBMA_FAMILY_GENERIC, HBMA_FAMILY_POOL, HBMA_FAMILY_WINDOW, HBMA_FAMILY_TILE = 0, 1, 2, 3, 4


class SvcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"svc_b200 error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    return os.environ.get("SVC_B200_LIB",
                          os.path.join(os.path.dirname(_PKG), "lib", "libsvc_b200.so"))


class _Cfg(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("frame_w", C.c_uint32), ("frame_h", C.c_uint32),
                ("mv_block_w", C.c_uint32), ("mv_block_h", C.c_uint32),
                ("mv_search_range", C.c_uint32), ("pyr_lvl_count", C.c_uint32),
                ("transform_block_w", C.c_uint32), ("transform_block_h", C.c_uint32),
                ("device", C.c_int32), ("max_batch", C.c_uint32), ("cuda_stream", C.c_void_p),
                ("hbma_kernel_family", C.c_uint32), ("host_chunk_frames", C.c_uint32)]


class _Info(C.Structure):
    _fields_ = [("padded_w", C.c_uint32), ("padded_h", C.c_uint32),
                ("mv_field_w", C.c_uint32), ("mv_field_h", C.c_uint32),
                ("frame_in_bytes", C.c_uint64), ("frame_stream_bytes", C.c_uint64),
                ("record_bytes", C.c_uint32), ("max_batch", C.c_uint32)]


_lib = None


def lib() -> C.CDLL:
    """Load libsvc_b200.so; a missing library is a hard error (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError(
            f"{p} not found: build it with `make -C scalable-video-codec_b200` "
            "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(p)
    L.svc_last_error.restype = C.c_char_p
    L.svc_version.restype = C.c_char_p
    L.svc_padded_dim.restype = C.c_uint32
    L.svc_padded_dim.argtypes = [C.c_uint32] * 3
    L.svc_serialized_frame_bytes.restype = C.c_uint64
    L.svc_serialized_frame_bytes.argtypes = [C.c_uint32] * 5
    L.svc_host_alloc.restype = C.c_void_p
    L.svc_host_alloc.argtypes = [C.c_size_t]
    L.svc_host_free.argtypes = [C.c_void_p]
    L.svc_device_alloc.restype = C.c_void_p
    L.svc_device_alloc.argtypes = [C.c_int, C.c_size_t]
    L.svc_device_free.argtypes = [C.c_int, C.c_void_p]
    L.svc_memcpy_h2d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.svc_memcpy_d2h.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.svc_session_create.argtypes = [C.POINTER(_Cfg), C.POINTER(C.c_void_p)]
    L.svc_session_destroy.argtypes = [C.c_void_p]
    L.svc_session_info_get.argtypes = [C.c_void_p, C.POINTER(_Info)]
    L.svc_session_reset.argtypes = [C.c_void_p]
    L.svc_session_synchronize.argtypes = [C.c_void_p]
    L.svc_session_launch_count.restype = C.c_uint64
    L.svc_session_launch_count.argtypes = [C.c_void_p]
    L.svc_session_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
    L.svc_session_encode_device.argtypes = L.svc_session_encode.argtypes
    L.svc_session_hbma_work.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.svc_session_run_stage.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise SvcError(rc, lib().svc_last_error().decode())


def _u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8:
        raise TypeError("expected uint8 data")
    return a


def _ptr_array(arrs: Sequence[np.ndarray], ty=_u8p):
    out = (ty * len(arrs))()
    for i, a in enumerate(arrs):
        out[i] = a.ctypes.data_as(ty)
    return out


def device_count() -> int:
    n = C.c_int(0)
    lib().svc_device_count(C.byref(n))
    return n.value


def sad_peak(device: int = 0) -> float:
    """Measured VABSDIFF4 peak of the device in byte-absdiffs per second."""
    v = C.c_double(0.0)
    _check(lib().svc_sad_peak(C.c_int(device), C.byref(v)))
    return v.value


def selftest_dequant(q_lo: int, q_hi: int, device: int = 0) -> int:
    """Mismatches between the decoder's division-free quotient and the IEEE division over every float
    in [-2^18, 2^18] and every quantisation step in [q_lo, q_hi] (must be 0)."""
    n = C.c_uint64(0)
    _check(lib().svc_selftest_dequant(C.c_int(device), C.c_uint32(q_lo), C.c_uint32(q_hi), C.byref(n)))
    if n.value:
        print(lib().svc_last_error().decode())
    return n.value


def padded_dim(a: int, mv_block: int, levels: int) -> int:
    return int(lib().svc_padded_dim(a, mv_block, levels))


def serialized_frame_bytes(w, h, tbw=8, tbh=8, channels=3) -> int:
    return int(lib().svc_serialized_frame_bytes(w, h, tbw, tbh, channels))


def write_header(n_input_frames, w, h, pw, ph, tbw=8, tbh=8, channels=3) -> np.ndarray:
    out = np.empty(32, np.uint8)
    _check(lib().svc_write_header(n_input_frames, w, h, pw, ph, tbw, tbh, channels,
                                  out.ctypes.data_as(_u8p)))
    return out


# ---- stateless drop-ins -------------------------------------------------------

def EstimateMotionHierarchical(tracked_pyramid, anchor_pyramid, level_count, frame_w, frame_h,
                               search_range, block_w, block_h):
    """libs/motion.hpp:134-138.  Returns (motion_field[mh,mw,2], min_mad[mh,mw])."""
    t = [_u8(x) for x in tracked_pyramid]
    a = [_u8(x) for x in anchor_pyramid]
    if len(t) < level_count or len(a) < level_count:
        raise ValueError("pyramids shorter than level_count")
    mw = frame_w // max(block_w, 1)
    mh = frame_h // max(block_h, 1)
    mv = np.empty((mh, mw, 2), np.float32)
    mad = np.empty((mh, mw), np.float32)
    _check(lib().svc_estimate_motion_hierarchical(
        _ptr_array(t), _ptr_array(a), C.c_uint32(level_count), C.c_uint32(frame_w),
        C.c_uint32(frame_h), C.c_uint32(search_range), C.c_uint32(block_w), C.c_uint32(block_h),
        mv.ctypes.data_as(_f32p), mad.ctypes.data_as(_f32p)))
    return mv, mad


def EstimateMotionHierarchical16x16Sse2(tracked_pyramid, anchor_pyramid, frame_w, frame_h,
                                        search_range):
    """libs/motion.hpp:148-152 (4 levels, 16x16 blocks)."""
    t = [_u8(x) for x in tracked_pyramid]
    a = [_u8(x) for x in anchor_pyramid]
    if len(t) < 4 or len(a) < 4:
        raise ValueError("pyramids need 4 levels")
    mv = np.empty((frame_h // 16, frame_w // 16, 2), np.float32)
    mad = np.empty((frame_h // 16, frame_w // 16), np.float32)
    _check(lib().svc_estimate_motion_hierarchical_16x16(
        _ptr_array(t), _ptr_array(a), C.c_uint32(frame_w), C.c_uint32(frame_h),
        C.c_uint32(search_range), mv.ctypes.data_as(_f32p), mad.ctypes.data_as(_f32p)))
    return mv, mad


def EstimateMotionExhaustiveSearch(tracked_frame, anchor_frame, frame_w, frame_h, search_range,
                                   block_w, block_h):
    """libs/motion.hpp:106-110."""
    t, a = _u8(tracked_frame), _u8(anchor_frame)
    mw = frame_w // max(block_w, 1)
    mh = frame_h // max(block_h, 1)
    mv = np.empty((mh, mw, 2), np.float32)
    mad = np.empty((mh, mw), np.float32)
    _check(lib().svc_estimate_motion_exhaustive(
        t.ctypes.data_as(_u8p), a.ctypes.data_as(_u8p), C.c_uint32(frame_w), C.c_uint32(frame_h),
        C.c_uint32(search_range), C.c_uint32(block_w), C.c_uint32(block_h),
        mv.ctypes.data_as(_f32p), mad.ctypes.data_as(_f32p)))
    return mv, mad


def y_pyramid(bgr, padded_w, padded_h, level_count) -> List[np.ndarray]:
    """copyMakeBorder + cvtColor(BGR2YUV)[0] + buildPyramid, libs/encoder.cpp:459-470."""
    bgr = _u8(bgr)
    h, w, _ = bgr.shape
    outs = [np.empty((padded_h >> l, padded_w >> l), np.uint8) for l in range(level_count)]
    _check(lib().svc_y_pyramid(bgr.ctypes.data_as(_u8p), C.c_uint32(w), C.c_uint32(h),
                               C.c_uint32(padded_w), C.c_uint32(padded_h),
                               C.c_uint32(level_count), _ptr_array(outs)))
    return outs


def dct_planar(bgr, padded_w, padded_h, tbw=8, tbh=8) -> np.ndarray:
    """convertTo(CV_32FC3) + Dct, libs/encoder.cpp:638-640 -> (3, ph, pw) float32 (B,G,R)."""
    bgr = _u8(bgr)
    h, w, _ = bgr.shape
    out = np.empty((3, padded_h, padded_w), np.float32)
    planes = (_f32p * 3)(*[out[c].ctypes.data_as(_f32p) for c in range(3)])
    _check(lib().svc_dct_planar(bgr.ctypes.data_as(_u8p), C.c_uint32(w), C.c_uint32(h),
                                C.c_uint32(padded_w), C.c_uint32(padded_h), C.c_uint32(tbw),
                                C.c_uint32(tbh), planes))
    return out


def encode_frame_stream(bgr, padded_w, padded_h, tbw=8, tbh=8, mv_block_w=16, mv_block_h=16,
                        block_types=None) -> np.ndarray:
    """Dct + SerializeEncodedFrame, libs/encoder.cpp:638-650 -> one frame's record bytes."""
    bgr = _u8(bgr)
    h, w, _ = bgr.shape
    out = np.empty(serialized_frame_bytes(w, h, tbw, tbh, 3), np.uint8)
    bt = None
    if block_types is not None:
        block_types = np.ascontiguousarray(block_types, dtype=np.uint32)
        bt = block_types.ctypes.data_as(_u32p)
    _check(lib().svc_encode_frame_stream(
        bgr.ctypes.data_as(_u8p), C.c_uint32(w), C.c_uint32(h), C.c_uint32(padded_w),
        C.c_uint32(padded_h), C.c_uint32(tbw), C.c_uint32(tbh), C.c_uint32(mv_block_w),
        C.c_uint32(mv_block_h), bt, out.ctypes.data_as(_u8p)))
    return out


def patch_block_types(frame_stream: np.ndarray, frame_w, frame_h, block_types, tbw=8, tbh=8,
                      channels=3, mv_block_w=16, mv_block_h=16, mv_field_w=None) -> None:
    block_types = np.ascontiguousarray(block_types, dtype=np.uint32)
    if mv_field_w is None:
        mv_field_w = block_types.shape[-1]
    assert frame_stream.dtype == np.uint8 and frame_stream.flags.c_contiguous
    _check(lib().svc_patch_block_types(
        frame_stream.ctypes.data_as(_u8p), C.c_uint32(frame_w), C.c_uint32(frame_h),
        C.c_uint32(tbw), C.c_uint32(tbh), C.c_uint32(channels), C.c_uint32(mv_block_w),
        C.c_uint32(mv_block_h), C.c_uint32(mv_field_w), block_types.ctypes.data_as(_u32p)))


class _Layout(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("frame_count", "frame_w", "frame_h", "padded_w", "padded_h", "tbw",
                                          "tbh", "channels", "record_bytes")] + \
               [("encoder_records_per_frame", C.c_uint64), ("decoder_records_per_frame", C.c_uint64),
                ("encoder_stream_bytes", C.c_uint64), ("consistent", C.c_int32)]


def stream_layout(header32) -> dict:
    """Record geometry implied by a stream header on the encoder and on the decoder side."""
    hdr = _u8(header32)
    out = _Layout()
    _check(lib().svc_stream_layout_from_header(hdr.ctypes.data_as(_u8p), C.byref(out)))
    return {n: getattr(out, n) for n, _ in _Layout._fields_}


class _Rect(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("w", C.c_uint32), ("h", C.c_uint32)]


def gaze_rect(gaze_x, gaze_y, max_w, max_h, frame_w, frame_h, padded_w, padded_h):
    """libs/decoder.cpp:66-98, 172-189 -> (x, y, w, h) in the padded frame."""
    r = _Rect()
    _check(lib().svc_gaze_rect(C.c_uint32(gaze_x), C.c_uint32(gaze_y), C.c_uint32(max_w), C.c_uint32(max_h),
                               C.c_uint32(frame_w), C.c_uint32(frame_h), C.c_uint32(padded_w),
                               C.c_uint32(padded_h), C.byref(r)))
    return (r.x, r.y, r.w, r.h)


def decode_frame_blocks(frame_records, padded_w, padded_h, fg_quant_step=1, bg_quant_step=640,
                        gaze=None, tbw=8, tbh=8) -> np.ndarray:
    """ParseBlock + DecodeBlock over one frame's records, libs/decoder.cpp:102-149, 191-213
    -> (padded_h, padded_w, 3) float32 BGR."""
    rec = _u8(frame_records)
    out = np.empty((padded_h, padded_w, 3), np.float32)
    g = C.byref(_Rect(*gaze)) if gaze is not None else None
    _check(lib().svc_decode_frame_blocks(rec.ctypes.data_as(_u8p), C.c_uint32(padded_w), C.c_uint32(padded_h),
                                         C.c_uint32(tbw), C.c_uint32(tbh), C.c_uint32(fg_quant_step),
                                         C.c_uint32(bg_quant_step), g, out.ctypes.data_as(_f32p)))
    return out


def decode_frames_device(device, cuda_stream, d_records, n_frames, padded_w, padded_h, d_out,
                         fg_quant_step=1, bg_quant_step=640, gaze=None, tb=8):
    g = C.byref(_Rect(*gaze)) if gaze is not None else None
    _check(lib().svc_decode_frames_device(C.c_int(device), C.c_void_p(cuda_stream or None),
                                          C.c_void_p(_addr(d_records)), C.c_uint32(n_frames),
                                          C.c_uint32(padded_w), C.c_uint32(padded_h), C.c_uint32(tb), C.c_uint32(tb),
                                          C.c_uint32(fg_quant_step), C.c_uint32(bg_quant_step), g,
                                          C.c_void_p(_addr(d_out))))


# ---- memory helpers -------------------------------------------------------------

class PinnedBuffer:
    """Pinned host memory from svc_host_alloc, exposed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self.ptr = lib().svc_host_alloc(self.nbytes)
        if not self.ptr:
            raise SvcError(2, f"svc_host_alloc({nbytes}) failed")
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.uint8)

    def view(self, dtype, shape):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return self.array[:n].view(dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().svc_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceBuffer:
    """Raw device memory from svc_device_alloc."""

    def __init__(self, device: int, nbytes: int):
        self.device, self.nbytes = device, int(nbytes)
        self.ptr = lib().svc_device_alloc(device, self.nbytes)
        if not self.ptr:
            raise SvcError(2, f"svc_device_alloc({nbytes}) failed")

    def upload(self, arr: np.ndarray, offset: int = 0):
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes
        _check(lib().svc_memcpy_h2d(self.device, self.ptr + offset, arr.ctypes.data, arr.nbytes))

    def download(self, dtype, shape, offset: int = 0) -> np.ndarray:
        out = np.empty(shape, dtype)
        assert offset + out.nbytes <= self.nbytes
        _check(lib().svc_memcpy_d2h(self.device, out.ctypes.data, self.ptr + offset, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().svc_device_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- session ----------------------------------------------------------------------

@dataclass
class SessionConfig:
    """VideoProperties + the hot-path fields of EncoderConfig (libs/encoder.hpp:25-50);
    defaults from apps/encoder.cpp:42-58."""
    frame_w: int
    frame_h: int
    mv_block_w: int = 16
    mv_block_h: int = 16
    mv_search_range: int = 8
    pyr_lvl_count: int = 4
    transform_block_w: int = 8
    transform_block_h: int = 8
    device: int = 0
    max_batch: int = 0
    cuda_stream: int = 0
    hbma_kernel_family: int = 0  # test hook: HBMA_FAMILY_*
    host_chunk_frames: int = 0   # frames per H2D | kernels | D2H pipeline stage of encode() (0 = 16)


def _addr(x) -> Optional[int]:
    if x is None:
        return None
    if isinstance(x, (PinnedBuffer, DeviceBuffer)):
        return x.ptr
    if isinstance(x, np.ndarray):
        assert x.flags.c_contiguous
        return x.ctypes.data
    return int(x)


class Session:
    """One Encoder (libs/encoder.hpp:52-95) on one GPU: device-resident hot path."""

    def __init__(self, cfg: SessionConfig):
        self.cfg = cfg
        c = _Cfg(C.sizeof(_Cfg), cfg.frame_w, cfg.frame_h, cfg.mv_block_w, cfg.mv_block_h,
                 cfg.mv_search_range, cfg.pyr_lvl_count, cfg.transform_block_w,
                 cfg.transform_block_h, cfg.device, cfg.max_batch, cfg.cuda_stream or None,
                 cfg.hbma_kernel_family, cfg.host_chunk_frames)
        h = C.c_void_p()
        _check(lib().svc_session_create(C.byref(c), C.byref(h)))
        self._h = h
        info = _Info()
        _check(lib().svc_session_info_get(self._h, C.byref(info)))
        self.padded_w, self.padded_h = info.padded_w, info.padded_h
        self.mv_field_w, self.mv_field_h = info.mv_field_w, info.mv_field_h
        self.frame_in_bytes = info.frame_in_bytes
        self.frame_stream_bytes = info.frame_stream_bytes
        self.record_bytes = info.record_bytes
        self.max_batch = info.max_batch

    def close(self):
        if self._h:
            lib().svc_session_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _check(lib().svc_session_reset(self._h))

    def synchronize(self):
        _check(lib().svc_session_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().svc_session_launch_count(self._h))

    def header(self, n_input_frames: int) -> np.ndarray:
        return write_header(n_input_frames, self.cfg.frame_w, self.cfg.frame_h, self.padded_w,
                            self.padded_h, self.cfg.transform_block_w,
                            self.cfg.transform_block_h, 3)

    def encode(self, frames_bgr: np.ndarray, want_mv=True, want_mad=True, want_stream=True,
               block_types=None, out_mv=None, out_mad=None, out_stream=None):
        """Host in / host out.  frames_bgr: (n, h, w, 3) uint8.  Returns
        (mv[n_enc,mh,mw,2], mad[n_enc,mh,mw], stream[n_enc,frame_stream_bytes])."""
        frames_bgr = _u8(frames_bgr)
        n = frames_bgr.shape[0]
        assert frames_bgr.shape[1:] == (self.cfg.frame_h, self.cfg.frame_w, 3)
        cap = n  # upper bound on encoded frames
        mh, mw = self.mv_field_h, self.mv_field_w
        mv = out_mv if out_mv is not None else (np.empty((cap, mh, mw, 2), np.float32) if want_mv else None)
        mad = out_mad if out_mad is not None else (np.empty((cap, mh, mw), np.float32) if want_mad else None)
        st = out_stream if out_stream is not None else (
            np.empty((cap, self.frame_stream_bytes), np.uint8) if want_stream else None)
        if block_types is not None:
            block_types = np.ascontiguousarray(block_types, dtype=np.uint32)
        ne = C.c_uint32(0)
        _check(lib().svc_session_encode(self._h, frames_bgr.ctypes.data, n, _addr(mv), _addr(mad),
                                        _addr(st), _addr(block_types), C.byref(ne)))
        k = ne.value
        return (mv[:k] if mv is not None else None, mad[:k] if mad is not None else None,
                st[:k] if st is not None else None)

    def encode_device(self, d_frames, n_frames, d_mv=None, d_mad=None, d_stream=None,
                      d_block_types=None) -> int:
        """Device pointers (int / DeviceBuffer); asynchronous on the session stream."""
        ne = C.c_uint32(0)
        _check(lib().svc_session_encode_device(self._h, _addr(d_frames), n_frames, _addr(d_mv),
                                               _addr(d_mad), _addr(d_stream),
                                               _addr(d_block_types), C.byref(ne)))
        return ne.value

    def hbma_work(self, n_frames):
        """(candidates, byte-absdiffs) K2 performs on pyramid slots 0..n_frames, exact."""
        c, a = C.c_uint64(0), C.c_uint64(0)
        _check(lib().svc_session_hbma_work(self._h, C.c_uint32(n_frames), C.byref(c), C.byref(a)))
        return c.value, a.value

    def run_stage(self, stage, d_frames, n_frames, d_mv=None, d_mad=None, d_stream=None):
        _check(lib().svc_session_run_stage(self._h, stage, _addr(d_frames), n_frames,
                                           _addr(d_mv), _addr(d_mad), _addr(d_stream)))
