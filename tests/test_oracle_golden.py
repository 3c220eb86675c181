"""CPU: the C oracle against the committed golden fixtures (compiled reference
motion code + cv2), i.e. the pin of the oracle itself."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden


@pytest.mark.parametrize("name,n", [("small_default.npz", 3), ("aligned_default.npz", 2)])
def test_y_pyramid_matches_cv2(oracle, name, n):
    g = load_golden(name)
    pw, ph = int(g["pw"]), int(g["ph"])
    for i in range(n):
        pyr = oracle.y_pyramid(g["frames"][i], pw, ph, 4)
        for l in range(4):
            assert np.array_equal(pyr[l], g[f"pyr{i}_{l}"]), (name, i, l)


def test_hbma_default_matches_reference_small(oracle):
    g = load_golden("small_default.npz")
    for i in (1, 2):
        t = [g[f"pyr{i-1}_{l}"] for l in range(4)]
        a = [g[f"pyr{i}_{l}"] for l in range(4)]
        mv, mad = oracle.hbma(t, a, 8)
        assert np.array_equal(mv, g[f"mv{i}"])
        assert np.array_equal(mad, g[f"mad{i}"])


@pytest.mark.parametrize("R", [8, 16, 32])
def test_hbma_default_matches_reference_aligned(oracle, R):
    g = load_golden("aligned_default.npz")
    t = [g[f"pyr0_{l}"] for l in range(4)]
    a = [g[f"pyr1_{l}"] for l in range(4)]
    mv, mad = oracle.hbma(t, a, R)
    assert np.array_equal(mv, g[f"mv_R{R}"])
    assert np.array_equal(mad, g[f"mad_R{R}"])


def test_hbma_generic_matches_reference(oracle):
    g = load_golden("generic_motion.npz")
    for ci, (lv, bw, bh, rr) in enumerate(g["cases"]):
        t = [g[f"t{ci}_{l}"] for l in range(lv)]
        a = [g[f"a{ci}_{l}"] for l in range(lv)]
        mv, mad = oracle.hbma(t, a, int(rr), int(bw), int(bh))
        assert np.array_equal(mv, g[f"mv{ci}"]), ci
        assert np.array_equal(mad, g[f"mad{ci}"]), ci
    mv, mad = oracle.ebma(g["t3"], g["a3"], 3, 8, 8)
    assert np.array_equal(mv, g["ebma_mv"]) and np.array_equal(mad, g["ebma_mad"])


def test_pyramid_levels_equal_pyrdown_of_golden(oracle):
    g = load_golden("generic_motion.npz")
    for ci, (lv, bw, bh, rr) in enumerate(g["cases"]):
        for l in range(1, lv):
            assert np.array_equal(oracle.pyr_down(g[f"t{ci}_{l-1}"]), g[f"t{ci}_{l}"])


# tolerance of the DCT contract: |coefficient error| <= 1e-3 absolute against
# OpenCV's float cv::dct (coefficients reach 2040; cv2 itself is ~1.2e-4 from
# the exact transform)
DCT_TOL = 1e-3


@pytest.mark.parametrize("name", ["small_default.npz", "aligned_default.npz"])
def test_dct_planes_within_tolerance_of_cv2(oracle, name):
    g = load_golden(name)
    d = oracle.dct_planar(g["frames"][1], int(g["pw"]), int(g["ph"]))
    assert np.abs(d - g["dct1"]).max() <= DCT_TOL


def test_dct_known_answers(oracle):
    g = load_golden("dct_kat.npz")
    for key, tbw, tbh in (("8", 8, 8), ("4", 4, 4), ("16", 16, 16), ("48", 8, 4)):
        for b, o in zip(g["b" + key], g["o" + key]):
            bgr = np.repeat(b.astype(np.uint8)[..., None], 3, axis=2)
            d = oracle.dct_planar(bgr, tbw, tbh, tbw, tbh)
            for c in range(3):
                assert np.abs(d[c] - o).max() <= DCT_TOL
    # DC of an all-255 block is 255 * 8 = 2040
    assert abs(g["o8"][0][0, 0] - 2040.0) < 1e-2


@pytest.mark.parametrize("name", ["small_default.npz", "aligned_default.npz"])
def test_serializer_matches_restated_reference_layout(oracle, name):
    g = load_golden(name)
    fr = g["frames"][1]
    h, w, _ = fr.shape
    pw = int(g["pw"])
    st = oracle.serialize_frame(g["dct1"], g["btypes"], w, h, 8, 8, pw // 16, 16, 16)
    assert st.size == oracle.serialized_frame_bytes(w, h)
    assert np.array_equal(np.frombuffer(hashlib.sha256(st.tobytes()).digest(), np.uint8),
                          g["stream1_sha256"])
    if "stream1_head" in g:
        assert np.array_equal(st[:g["stream1_head"].size], g["stream1_head"])


def test_header_layout(oracle):
    hdr = oracle.header(300, 1920, 1080, 1920, 1088).view(np.uint32)
    assert list(hdr) == [299, 1920, 1080, 0, 8, 8, 8, 3]


def test_padding_rule(oracle):
    # libs/encoder.cpp:165-169 with 16x16 blocks and 4 levels
    assert oracle.padded_dim(1080, 16, 4) == 1088
    assert oracle.padded_dim(540, 16, 4) == 544
    assert oracle.padded_dim(2160, 16, 4) == 2160
    assert oracle.padded_dim(100, 8, 5) == 112
