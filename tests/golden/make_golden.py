"""Generates the committed golden fixtures under tests/golden/.

Run in the authoring container (needs /root/reference compiled into
oracle/_ref by `make -C oracle`, and python cv2):

    python tests/golden/make_golden.py

Sources of truth recorded here (the reference repository has no tests or
golden vectors of its own):
  * motion vectors / MADs: the UNMODIFIED reference libs/motion.cpp
    (EstimateMotionHierarchical16x16Sse2, EstimateMotionHierarchical,
    EstimateMotionExhaustiveSearch) through oracle/_ref/libref_motion.so;
  * Y plane, pyramid, DCT: OpenCV (python cv2) -- the same calls the reference
    makes at libs/encoder.cpp:447-451, 459-470, 323-339, 638;
  * stream bytes: a direct python restatement of SerializeEncodedFrame
    (libs/encoder.cpp:243-266) applied to the cv2 coefficient planes.
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scalable-video-codec_b200"))

from oracle import oracle as O  # noqa: E402
from svc_b200.synth import SyntheticSequence  # noqa: E402


def lcm_pad(a, block, levels):
    l = np.lcm(block, 1 << (levels - 1))
    return int((a + l - 1) // l * l)


def cv_pyramid(bgr, pw, ph, levels):
    h, w, _ = bgr.shape
    p = cv2.copyMakeBorder(bgr, 0, ph - h, 0, pw - w, cv2.BORDER_CONSTANT, value=(0, 0, 0))
    y = np.ascontiguousarray(cv2.cvtColor(p, cv2.COLOR_BGR2YUV)[..., 0])
    out = [y]
    for _ in range(levels - 1):
        out.append(cv2.pyrDown(out[-1]))
    return out


def cv_dct_planes(bgr, pw, ph, tbw, tbh):
    h, w, _ = bgr.shape
    p = cv2.copyMakeBorder(bgr, 0, ph - h, 0, pw - w, cv2.BORDER_CONSTANT, value=(0, 0, 0))
    f = p.astype(np.float32)
    planes = [np.ascontiguousarray(f[..., c]) for c in range(3)]
    for pl in planes:
        for y in range(0, ph, tbh):
            for x in range(0, pw, tbw):
                pl[y:y + tbh, x:x + tbw] = cv2.dct(pl[y:y + tbh, x:x + tbw])
    return np.stack(planes)


def py_serialize(planes, btypes, w, h, tbw, tbh, mvw, mbw, mbh):
    """SerializeEncodedFrame, libs/encoder.cpp:243-266 (unpadded w/h; flat index
    with the unpadded width; tbw rows of tbh floats)."""
    flat = [pl.ravel() for pl in planes]
    out = bytearray()
    for tb_y in range(0, h, tbh):
        for tb_x in range(0, w, tbw):
            bt = np.uint32(btypes[(tb_y // mbh) * mvw + tb_x // mbw])
            out += bt.tobytes()
            for ch in flat:
                for y in range(tb_y, tb_y + tbw):
                    s = y * w + tb_x
                    out += ch[s:s + tbh].astype(np.float32).tobytes()
    return np.frombuffer(bytes(out), np.uint8)


def make_decode_golden():
    """Decoder block path (libs/decoder.cpp:102-149, 191-213) restated with numpy + cv2.idct:
    records -> quantise (round half away from zero, float) -> cv2.idct -> merge."""
    rng = np.random.default_rng(77)
    pw, ph = 80, 48  # 10 x 6 blocks: a partial 32-block chunk
    f = rng.integers(0, 256, (ph, pw, 3)).astype(np.uint8)
    planes = cv_dct_planes(f, pw, ph, 8, 8)
    bt = rng.integers(0, 3, (ph // 16) * (pw // 16)).astype(np.uint32)
    st = py_serialize(planes, bt, pw, ph, 8, 8, pw // 16, 16, 16)
    rec = st.view(np.uint32).reshape(-1, 193)
    outs = {}
    for name, fg, bg, gaze in (("a", 1, 640, None), ("b", 3, 40, (8, 8, 32, 16)), ("c", 7, 7, (0, 0, 80, 48))):
        exp = np.zeros((ph, pw, 3), np.float32)
        k = 0
        for y in range(0, ph, 8):
            for x in range(0, pw, 8):
                t = rec[k, 0]
                gazed = gaze is not None and gaze[0] <= x < gaze[0] + gaze[2] and gaze[1] <= y < gaze[1] + gaze[3]
                q = np.float32(1 if gazed else (bg if t == 0 else fg))
                for c in range(3):
                    b = rec[k, 1 + 64 * c:65 + 64 * c].view(np.float32).reshape(8, 8) / q
                    b = np.where(b >= 0, np.floor(b + np.float32(0.5)), np.ceil(b - np.float32(0.5))).astype(np.float32) * q
                    exp[y:y + 8, x:x + 8, c] = cv2.idct(b)
                k += 1
        outs["out_" + name] = exp
        outs["cfg_" + name] = np.array([fg, bg] + (list(gaze) if gaze else [0, 0, 0, 0]) + [int(gaze is not None)], np.int64)
    np.savez_compressed(os.path.join(HERE, "decode_blocks.npz"), records=st, pw=pw, ph=ph, **outs)


def make_decode_golden_square(tb):
    """The same for 16x16 / 4x4 transform blocks (tests/golden/decode_blocks_tb{tb}.npz): cv2.dct
    planes -> records -> quantise -> cv2.idct -> merge."""
    rng = np.random.default_rng(770 + tb)
    pw, ph = 144, 48  # 9 x 3 blocks of 16: a partial 8-record unit; 36 x 12 blocks of 4
    f = rng.integers(0, 256, (ph, pw, 3)).astype(np.uint8)
    planes = cv_dct_planes(f, pw, ph, tb, tb)
    bt = rng.integers(0, 3, (ph // 16) * (pw // 16)).astype(np.uint32)
    st = py_serialize(planes, bt, pw, ph, tb, tb, pw // 16, 16, 16)
    a = tb * tb
    rec = st.view(np.uint32).reshape(-1, 1 + 3 * a)
    outs = {}
    for name, fg, bg, gaze in (("a", 1, 640, None), ("b", 3, 40, (16, 16, 64, 16)), ("c", 5, 5, (0, 0, pw, ph))):
        exp = np.zeros((ph, pw, 3), np.float32)
        k = 0
        for y in range(0, ph, tb):
            for x in range(0, pw, tb):
                t = rec[k, 0]
                gazed = gaze is not None and gaze[0] <= x < gaze[0] + gaze[2] and gaze[1] <= y < gaze[1] + gaze[3]
                q = np.float32(1 if gazed else (bg if t == 0 else fg))
                for c in range(3):
                    b = rec[k, 1 + a * c:1 + a * (c + 1)].view(np.float32).reshape(tb, tb) / q
                    b = np.where(b >= 0, np.floor(b + np.float32(0.5)), np.ceil(b - np.float32(0.5))).astype(np.float32) * q
                    exp[y:y + tb, x:x + tb, c] = cv2.idct(b)
                k += 1
        outs["out_" + name] = exp
        outs["cfg_" + name] = np.array([fg, bg] + (list(gaze) if gaze else [0, 0, 0, 0]) + [int(gaze is not None)], np.int64)
    np.savez_compressed(os.path.join(HERE, "decode_blocks_tb%d.npz" % tb), records=st, pw=pw, ph=ph, tb=tb, **outs)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "decode":
        make_decode_golden()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "decode_square":
        make_decode_golden_square(16)
        make_decode_golden_square(4)
        return
    assert O.have_ref(), "build oracle/_ref first: make -C oracle"
    rng = np.random.default_rng(20260101)

    # ---- case "small": W != padded W, default config, full hot path -----------
    w, h, L, B, R = 104, 56, 4, 16, 8
    pw, ph = lcm_pad(w, B, L), lcm_pad(h, B, L)
    seq = SyntheticSequence(w, h, 3, seed=7, n_rects=3)
    fr = seq.frames()
    pyr = [cv_pyramid(f, pw, ph, L) for f in fr]
    d = {"frames": fr, "pw": pw, "ph": ph}
    for i in range(3):
        for l in range(L):
            d[f"pyr{i}_{l}"] = pyr[i][l]
    for i in (1, 2):
        mv, mad = O.hbma(pyr[i - 1], pyr[i], R, impl="ref_sse2")
        mvg, madg = O.hbma(pyr[i - 1], pyr[i], R, impl="ref")
        assert np.array_equal(mv, mvg) and np.array_equal(mad, madg)
        d[f"mv{i}"], d[f"mad{i}"] = mv, mad
    planes = cv_dct_planes(fr[1], pw, ph, 8, 8)
    d["dct1"] = planes
    mvw = pw // B
    bt = rng.integers(0, 5, size=(ph // B) * mvw).astype(np.uint32)
    d["btypes"] = bt
    st = py_serialize(planes, bt, w, h, 8, 8, mvw, B, B)
    d["stream1_sha256"] = np.frombuffer(hashlib.sha256(st.tobytes()).digest(), np.uint8)
    d["stream1_head"] = st[:772 * 3]
    np.savez_compressed(os.path.join(HERE, "small_default.npz"), **d)

    # ---- case "aligned": W == padded W, H not a multiple of 8 ------------------
    w, h = 160, 92
    pw, ph = lcm_pad(w, B, L), lcm_pad(h, B, L)
    seq = SyntheticSequence(w, h, 2, seed=11, n_rects=4)
    fr = seq.frames()
    pyr = [cv_pyramid(f, pw, ph, L) for f in fr]
    d = {"frames": fr, "pw": pw, "ph": ph}
    for i in range(2):
        for l in range(L):
            d[f"pyr{i}_{l}"] = pyr[i][l]
    for R in (8, 16, 32):
        d[f"mv_R{R}"], d[f"mad_R{R}"] = O.hbma(pyr[0], pyr[1], R, impl="ref_sse2")
    planes = cv_dct_planes(fr[1], pw, ph, 8, 8)
    d["dct1"] = planes
    mvw = pw // B
    bt = rng.integers(0, 9, size=(ph // B) * mvw).astype(np.uint32)
    d["btypes"] = bt
    st = py_serialize(planes, bt, w, h, 8, 8, mvw, B, B)
    d["stream1_sha256"] = np.frombuffer(hashlib.sha256(st.tobytes()).digest(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "aligned_default.npz"), **d)

    # ---- generic-signature sweep on raw luminance pyramids ----------------------
    # (levels, block_w, block_h, range); tracked/anchor built with cv2.pyrDown
    d = {}
    cases = [(1, 16, 16, 4), (2, 16, 16, 6), (3, 8, 8, 8), (4, 16, 16, 24), (5, 16, 16, 16),
             (3, 16, 8, 5), (2, 12, 6, 4), (1, 5, 3, 2)]
    d["cases"] = np.array(cases, np.int32)
    for ci, (lv, bw, bh, rr) in enumerate(cases):
        fw = bw * 9
        fh = bh * 7
        base = rng.integers(0, 256, size=(fh + 16, fw + 16)).astype(np.uint8)
        base = cv2.GaussianBlur(base, (0, 0), 1.5)
        t0 = np.ascontiguousarray(base[8:8 + fh, 8:8 + fw])
        a0 = np.ascontiguousarray(base[8 + 2:8 + 2 + fh, 8 - 3:8 - 3 + fw])
        a0 = np.clip(a0.astype(np.int16) + rng.integers(-2, 3, a0.shape), 0, 255).astype(np.uint8)
        a0[: bh * 2, : bw * 2] = 77  # flat corner: ties
        t0[: bh * 2, : bw * 2] = 77
        tp, ap = [t0], [a0]
        for _ in range(lv - 1):
            tp.append(cv2.pyrDown(tp[-1]))
            ap.append(cv2.pyrDown(ap[-1]))
        mv, mad = O.hbma(tp, ap, rr, bw, bh, impl="ref")
        d[f"t{ci}"], d[f"a{ci}"] = t0, a0
        for l in range(lv):
            d[f"t{ci}_{l}"], d[f"a{ci}_{l}"] = tp[l], ap[l]
        d[f"mv{ci}"], d[f"mad{ci}"] = mv, mad
    # EBMA directly
    t0, a0 = d["t3"], d["a3"]
    d["ebma_mv"], d["ebma_mad"] = O.ebma(t0, a0, 3, 8, 8, impl="ref")
    np.savez_compressed(os.path.join(HERE, "generic_motion.npz"), **d)

    # ---- DCT known answers (cv2.dct on single blocks) ----------------------------
    blocks = rng.integers(0, 256, size=(16, 8, 8)).astype(np.float32)
    blocks[0] = 255.0
    blocks[1] = 0.0
    blocks[2] = np.arange(64, dtype=np.float32).reshape(8, 8)
    out = np.stack([cv2.dct(b) for b in blocks])
    b4 = rng.integers(0, 256, size=(4, 4, 4)).astype(np.float32)
    o4 = np.stack([cv2.dct(b) for b in b4])
    b16 = rng.integers(0, 256, size=(2, 16, 16)).astype(np.float32)
    o16 = np.stack([cv2.dct(b) for b in b16])
    b48 = rng.integers(0, 256, size=(3, 4, 8)).astype(np.float32)  # 4 rows x 8 cols
    o48 = np.stack([cv2.dct(b) for b in b48])
    np.savez_compressed(os.path.join(HERE, "dct_kat.npz"), b8=blocks, o8=out, b4=b4, o4=o4,
                        b16=b16, o16=o16, b48=b48, o48=o48)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
