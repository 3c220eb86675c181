"""Decoder block path: dequantise + 8x8 IDCT + merge.
CPU: oracle vs the cv2.idct golden fixture, gaze rectangle arithmetic.
GPU: CUDA kernel vs oracle / golden on the same record bytes, encode -> decode round trip."""
import numpy as np
import pytest

from conftest import load_golden

IDCT_TOL = 1e-3  # absolute, on pixel-scale values (cv2.idct itself is ~5e-5 from the exact inverse)


def _cases(g):
    for name in "abc":
        cfg = g["cfg_" + name]
        gaze = tuple(int(v) for v in cfg[2:6]) if cfg[6] else None
        yield int(cfg[0]), int(cfg[1]), gaze, g["out_" + name]


def test_oracle_decode_matches_cv2_golden(oracle):
    g = load_golden("decode_blocks.npz")
    for fg, bg, gaze, exp in _cases(g):
        out = oracle.decode_frame_blocks(g["records"], int(g["pw"]), int(g["ph"]), fg_q=fg, bg_q=bg, gaze=gaze)
        assert np.abs(out - exp).max() <= IDCT_TOL


@pytest.mark.parametrize("tb", [16, 4])
def test_oracle_decode_square_blocks_matches_cv2_golden(oracle, tb):
    """16x16 / 4x4 records: the oracle against cv2.idct (tests/golden/make_golden.py decode_square)."""
    g = load_golden("decode_blocks_tb%d.npz" % tb)
    for fg, bg, gaze, exp in _cases(g):
        out = oracle.decode_frame_blocks(g["records"], int(g["pw"]), int(g["ph"]), tb, tb, fg_q=fg, bg_q=bg, gaze=gaze)
        assert np.abs(out - exp).max() <= IDCT_TOL


@pytest.mark.parametrize("args", [(100, 50, 64, 64, 960, 540, 960, 544), (5, 535, 64, 64, 960, 540, 960, 544),
                                  (0, 0, 64, 64, 1920, 1080, 1920, 1088), (1919, 1079, 64, 64, 1920, 1080, 1920, 1088),
                                  (500, 300, 31, 77, 1000, 600, 1008, 608)])
def test_gaze_rect_matches_oracle(svc, oracle, args):
    assert svc.gaze_rect(*args) == oracle.gaze_rect(*args)


def test_decode_argument_validation(svc):
    rec = np.zeros(772 * 4, np.uint8)
    out_shape = (16, 16)
    with pytest.raises(svc.SvcError) as e:
        svc.decode_frame_blocks(rec, 16, 16, fg_quant_step=0)
    assert e.value.code == 1 and "foreground quantization step" in str(e.value)
    with pytest.raises(svc.SvcError) as e:
        svc.decode_frame_blocks(rec, 16, 16, bg_quant_step=0)
    assert "background quantization step" in str(e.value)
    with pytest.raises(svc.SvcError) as e:
        svc.decode_frame_blocks(rec, 16, 16, tbw=8, tbh=4)
    assert e.value.code == 3
    with pytest.raises(svc.SvcError) as e:
        svc.decode_frame_blocks(rec, 24, 16, tbw=16, tbh=16)  # 24 is not a multiple of 16
    assert e.value.code == 1


@pytest.mark.gpu
def test_gpu_decode_golden(gpu):
    g = load_golden("decode_blocks.npz")
    for fg, bg, gaze, exp in _cases(g):
        out = gpu.decode_frame_blocks(g["records"], int(g["pw"]), int(g["ph"]), fg, bg, gaze)
        assert np.abs(out - exp).max() <= IDCT_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("tb", [16, 4])
def test_gpu_decode_square_blocks_golden(gpu, tb):
    g = load_golden("decode_blocks_tb%d.npz" % tb)
    for fg, bg, gaze, exp in _cases(g):
        out = gpu.decode_frame_blocks(g["records"], int(g["pw"]), int(g["ph"]), fg, bg, gaze, tbw=tb, tbh=tb)
        assert np.abs(out - exp).max() <= IDCT_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("pw,ph", [(1920, 1088), (960, 544), (264, 40), (8, 8), (3840, 2160)])
@pytest.mark.parametrize("fg,bg,with_gaze", [(1, 640, False), (2, 37, True)])
def test_gpu_decode_vs_oracle(gpu, oracle, pw, ph, fg, bg, with_gaze):
    rng = np.random.default_rng(pw + ph + fg)
    n = (pw // 8) * (ph // 8)
    rec = np.empty((n, 193), np.uint32)
    rec[:, 0] = rng.integers(0, 3, n)
    coef = (rng.standard_normal((n, 192)) * 300).astype(np.float32)
    coef[:, ::64] = rng.uniform(0, 2040, (n, 3)).astype(np.float32)  # DC terms
    rec[:, 1:] = coef.view(np.uint32)
    gaze = (pw // 4 // 8 * 8, 0, min(64, pw), min(48, ph)) if with_gaze else None
    got = gpu.decode_frame_blocks(rec.view(np.uint8).ravel(), pw, ph, fg, bg, gaze)
    if pw * ph > 1000 * 1000:  # the oracle on a crop of whole block rows (records are row-major)
        rows = 16
        exp = oracle.decode_frame_blocks(rec[: (pw // 8) * (rows // 8)].view(np.uint8).ravel(), pw, rows,
                                         fg_q=fg, bg_q=bg, gaze=gaze)
        assert np.abs(got[:rows] - exp).max() <= IDCT_TOL
        tail = rec[-(pw // 8) * 2:]
        gz = None if gaze is None else (gaze[0], 0, 0, 0)  # the last rows are outside the gaze rectangle
        exp = oracle.decode_frame_blocks(tail.view(np.uint8).ravel(), pw, 16, fg_q=fg, bg_q=bg, gaze=gz)
        assert np.abs(got[-16:] - exp).max() <= IDCT_TOL
    else:
        exp = oracle.decode_frame_blocks(rec.view(np.uint8).ravel(), pw, ph, fg_q=fg, bg_q=bg, gaze=gaze)
        assert np.abs(got - exp).max() <= IDCT_TOL


@pytest.mark.gpu
def test_gpu_encode_decode_round_trip(gpu):
    """Stream written by K3 for a frame that needs no padding, decoded with q = 1 (which still
    rounds every coefficient to an integer, libs/decoder.cpp:140-142): the pixels return within
    the rounding noise of 64 orthonormal coefficients (rms ~0.29, bounded by 0.5 * 8)."""
    from svc_b200.synth import SyntheticSequence
    w, h = 320, 176  # multiples of 16: padded == unpadded, encoder and decoder record counts agree
    f = SyntheticSequence(w, h, 2, seed=5).frames()
    st = gpu.encode_frame_stream(f[1], w, h)
    dec = gpu.decode_frame_blocks(st, w, h, fg_quant_step=1, bg_quant_step=1)
    err = np.abs(dec - f[1].astype(np.float32))
    assert err.max() <= 4.0 and err.mean() < 0.4
    # coarse background quantisation: still the same picture within the quantisation error
    dec = gpu.decode_frame_blocks(st, w, h, fg_quant_step=1, bg_quant_step=16)
    assert np.abs(dec - f[1].astype(np.float32)).mean() < 8.0


@pytest.mark.gpu
@pytest.mark.parametrize("tb", [4, 16])
@pytest.mark.parametrize("pw,ph", [(1920, 1088), (272, 48), (16, 16), (80, 32), (2064, 32)])
@pytest.mark.parametrize("fg,bg,with_gaze", [(1, 640, False), (3, 29, True)])
def test_gpu_decode_square_blocks_vs_oracle(gpu, oracle, tb, pw, ph, fg, bg, with_gaze):
    """16x16 and 4x4 records (the streams of the fused dct16x16 / dct4x4 kernels) through
    idct16x16_decode_kernel / idct4x4_decode_kernel against the oracle's ParseBlock + DecodeBlock."""
    rng = np.random.default_rng(pw + ph + fg + tb)
    n, a = (pw // tb) * (ph // tb), tb * tb
    rec = np.empty((n, 1 + 3 * a), np.uint32)
    rec[:, 0] = rng.integers(0, 3, n)
    coef = (rng.standard_normal((n, 3 * a)) * 300).astype(np.float32)
    coef[:, ::a] = rng.uniform(0, 255 * tb, (n, 3)).astype(np.float32)  # DC terms
    rec[:, 1:] = coef.view(np.uint32)
    gaze = (pw // 4 // tb * tb, 0, min(64, pw), min(48, ph)) if with_gaze else None
    got = gpu.decode_frame_blocks(rec.view(np.uint8).ravel(), pw, ph, fg, bg, gaze, tbw=tb, tbh=tb)
    rows = 32 if pw * ph > 1000 * 1000 else ph   # the oracle on the first block rows of a large frame
    exp = oracle.decode_frame_blocks(rec[: (pw // tb) * (rows // tb)].view(np.uint8).ravel(), pw, rows, tb, tb,
                                     fg_q=fg, bg_q=bg, gaze=gaze)
    assert np.abs(got[:rows] - exp).max() <= IDCT_TOL
    if rows < ph:
        tail = rec[-(pw // tb) * (32 // tb):]
        gz = None if gaze is None else (gaze[0], 0, 0, 0)
        exp = oracle.decode_frame_blocks(tail.view(np.uint8).ravel(), pw, 32, tb, tb, fg_q=fg, bg_q=bg, gaze=gz)
        assert np.abs(got[-32:] - exp).max() <= IDCT_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("tb", [4, 16])
def test_gpu_encode_decode_round_trip_square_blocks(gpu, tb):
    from svc_b200.synth import SyntheticSequence
    w, h = 320, 176
    f = SyntheticSequence(w, h, 2, seed=5 + tb).frames()
    st = gpu.encode_frame_stream(f[1], w, h, tb, tb)
    dec = gpu.decode_frame_blocks(st, w, h, fg_quant_step=1, bg_quant_step=1, tbw=tb, tbh=tb)
    err = np.abs(dec - f[1].astype(np.float32))
    assert err.max() <= 0.5 * tb + 0.5 and err.mean() < 0.4


@pytest.mark.gpu
def test_dequantiser_quotient_is_the_ieee_division_exhaustively(gpu):
    """The decoder kernels compute round(c / q) * q (libs/decoder.cpp:137-144) without the IEEE
    division subroutine (reciprocal + two FMAs, k_idct.cu).  That quotient must be bit-identical
    to the division for every float c in [-2^18, 2^18] and every step the fast path accepts."""
    assert gpu.selftest_dequant(1, 4096) == 0
    assert gpu.selftest_dequant(4097, 4100) == 0  # beyond the fast-path limit: the IEEE division itself
