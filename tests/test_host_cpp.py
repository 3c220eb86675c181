"""The C++ host layer (scalable-video-codec_b200/host): reference-compatible
motion.hpp signatures and the Encoder functor, checked by tests/cpp/test_host.cpp
against the C oracle."""
import os
import subprocess

import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "scalable-video-codec_b200")
BIN = os.path.join(PKG, "build", "test_host")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True,
                   stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", PKG, "build/test_host"], check=True, stdout=subprocess.DEVNULL)


def test_host_layer_builds_and_validates_without_gpu():
    _build()
    r = subprocess.run([BIN, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_layer_encoder_matches_oracle(gpu):
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS frames=6" in r.stdout, r.stdout + r.stderr


ENC = os.path.join(PKG, "bin", "svc_encoder")


def _build_app():
    subprocess.run(["make", "-C", PKG, "bin/svc_encoder"], check=True, stdout=subprocess.DEVNULL)


def test_encoder_app_rejects_bad_config_like_the_reference(tmp_path):
    _build_app()
    raw = tmp_path / "in.bgr"
    raw.write_bytes(bytes(64 * 48 * 3 * 2))
    r = subprocess.run([ENC, "--width", "64", "--height", "48", "--mv-search-range", "4", str(raw)],
                       capture_output=True, timeout=60)
    assert r.returncode != 0
    assert b"Invalid encoder configuration" in r.stderr and b"quotient" in r.stderr
    assert subprocess.run([ENC], capture_output=True, timeout=60).returncode != 0  # usage


@pytest.mark.gpu
def test_encoder_app_stream_matches_oracle(gpu, oracle, tmp_path):
    import numpy as np
    from svc_b200.synth import SyntheticSequence
    _build_app()
    w, h, n = 320, 180, 6
    frames = SyntheticSequence(w, h, n, seed=31).frames()
    raw = tmp_path / "in.bgr"
    raw.write_bytes(frames.tobytes())
    r = subprocess.run([ENC, "--width", str(w), "--height", str(h), "--batch", "4", "--verbose", "0",
                        "--segment", "0", str(raw)], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.frombuffer(r.stdout, np.uint8)
    pw, ph = oracle.padded_dim(w, 16, 4), oracle.padded_dim(h, 16, 4)
    fb = oracle.serialized_frame_bytes(w, h)
    assert out.size == 32 + (n - 1) * fb
    assert np.array_equal(out[:32], oracle.header(n, w, h, pw, ph))
    for i in range(1, n):
        exp = oracle.serialize_frame(oracle.dct_planar(frames[i], pw, ph), None, w, h, 8, 8, pw // 16, 16, 16)
        got = out[32 + (i - 1) * fb: 32 + i * fb]
        g = got.view(np.uint32).reshape(-1, 193)
        e = exp.view(np.uint32).reshape(-1, 193)
        assert not g[:, 0].any()
        assert np.abs(g[:, 1:].view(np.float32) - e[:, 1:].view(np.float32)).max() <= 1e-3


@pytest.mark.gpu
def test_encoder_app_sharded_output_is_identical_to_single_device(gpu, tmp_path):
    """Frame-range sharding over several sessions (here: the same GPU listed 1, 2 and 3 times)
    must give byte-identical streams, and the same bytes as the streaming (stdout) mode."""
    import numpy as np
    from svc_b200.synth import SyntheticSequence
    _build_app()
    w, h, n = 320, 180, 11
    raw = tmp_path / "in.bgr"
    raw.write_bytes(SyntheticSequence(w, h, n, seed=41).frames().tobytes())
    outs = []
    for devs in ("0", "0,0", "0,0,0"):
        out = tmp_path / ("out_%d.svc" % len(devs))
        r = subprocess.run([ENC, "--width", str(w), "--height", str(h), "--batch", "4", "--verbose", "0",
                            "--seed", "7", "--devices", devs, "--out", str(out), str(raw)],
                           capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(out.read_bytes())
    r = subprocess.run([ENC, "--width", str(w), "--height", str(h), "--batch", "3", "--verbose", "0",
                        "--seed", "7", "--classify-threads", "3", str(raw)], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert np.frombuffer(outs[0], np.uint8)[32:].view(np.uint32).reshape(-1, 193)[:, 0].any()  # real block types
    assert len(outs[0]) == 32 + (n - 1) * gpu.serialized_frame_bytes(w, h)
    assert outs[0] == outs[1] == outs[2] == r.stdout


@pytest.mark.gpu
def test_encoder_app_block_types_follow_the_reference_chain(gpu, oracle, tmp_path):
    """The application's stream carries the block types of libs/encoder.cpp:491-624 computed from the
    GPU motion field: RANSAC (compiled reference), morphology / k-means / connected components
    (python cv2) on the ORACLE's motion field, with the per-frame generator states of --seed."""
    import numpy as np
    cv2 = pytest.importorskip("cv2")  # noqa: F841
    from svc_b200 import segment as seg
    from svc_b200.synth import SyntheticSequence
    _build_app()
    w, h, n, seed = 640, 368, 5, 5
    frames = SyntheticSequence(w, h, n, seed=77).frames()
    raw = tmp_path / "in.bgr"
    raw.write_bytes(frames.tobytes())
    r = subprocess.run([ENC, "--width", str(w), "--height", str(h), "--batch", "3", "--verbose", "0",
                        "--seed", str(seed), "--kmeans-cluster-count", "4", "--ransac-inlier-thresh", "2.5", str(raw)],
                       capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.frombuffer(r.stdout, np.uint8)
    pw, ph = oracle.padded_dim(w, 16, 4), oracle.padded_dim(h, 16, 4)
    mw, mh = pw // 16, ph // 16
    fb = oracle.serialized_frame_bytes(w, h)
    assert out.size == 32 + (n - 1) * fb
    pyr = [oracle.y_pyramid(frames[i], pw, ph, 4) for i in range(n)]
    any_fg = False
    for t in range(1, n):
        mv, _ = oracle.hbma(pyr[t - 1], pyr[t], 8)
        rs, ks = seg.frame_generators(seed, t - 1)
        L = oracle.ref_seeded(rs)
        if L is None:
            pytest.skip("oracle/_ref not built")
        _, _, inl = oracle.ref_ransac(L, mv, 1, 2.5, 0.99, 0.5)
        exp = oracle.block_types_cv2(mv, inl, ks, cluster_count=4).reshape(-1)
        assert np.array_equal(seg.block_types(mv, seg.SegmentConfig(kmeans_cluster_count=4, ransac_inlier_thresh=2.5),
                                              rs, ks)[0].reshape(-1), exp)
        got = out[32 + (t - 1) * fb: 32 + t * fb].view(np.uint32).reshape(-1, 193)[:, 0]
        by, bx = np.divmod(np.arange(got.size), w // 8)   # record order: transform blocks, raster
        assert np.array_equal(got, exp[(by * 8 // 16) * mw + bx * 8 // 16])
        any_fg |= bool(got.any())
    assert any_fg


DEC = os.path.join(PKG, "bin", "svc_decoder")


def _build_decoder():
    subprocess.run(["make", "-C", PKG, "bin/svc_decoder"], check=True, stdout=subprocess.DEVNULL)


def test_decoder_app_validates_like_the_reference(oracle):
    """Validate(DecoderConfig) messages (libs/decoder.cpp:35-47), short header, and the streams this
    decoder refuses: other transform blocks, horizontally padded frames."""
    _build_decoder()
    run = lambda args, data=b"": subprocess.run([DEC] + args, input=data, capture_output=True, timeout=60)
    assert run([]).returncode != 0
    r = run(["--foreground-quant-step", "0", "-"])
    assert r.returncode != 0 and b"invalid foreground quantization step" in r.stderr
    r = run(["--background-quant-step", "0", "-"])
    assert r.returncode != 0 and b"invalid background quantization step" in r.stderr
    r = run(["-"], b"short")
    assert r.returncode != 0 and b"header" in r.stderr
    hdr = oracle.header(3, 1000, 600, 1008, 608).tobytes()  # 1000 -> 1008: horizontal padding
    r = run(["-"], hdr)
    assert r.returncode != 0 and b"not decodable" in r.stderr
    hdr = oracle.header(3, 64, 64, 64, 64, tbw=8, tbh=4).tobytes()
    r = run(["-"], hdr)
    assert r.returncode != 0 and b"8x8" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,gaze,tb", [(320, 176, None, 8), (320, 180, (100, 60), 8), (1920, 1080, (960, 1070), 8),
                                         (320, 180, (100, 60), 16), (320, 172, None, 4)])
def test_encoder_app_piped_into_decoder_app(gpu, oracle, tmp_path, w, h, gaze, tb):
    """svc_encoder | svc_decoder: every decoded frame equals the oracle's ParseBlock + DecodeBlock
    (libs/decoder.cpp:102-149) of the same records, rounded to 8 bits; at 1080p / 180 rows the encoder
    writes fewer block rows than the padded frame holds and the rest stays black."""
    import numpy as np
    from svc_b200.synth import SyntheticSequence
    _build_app()
    _build_decoder()
    n = 4
    frames = SyntheticSequence(w, h, n, seed=w + h).frames()
    raw = tmp_path / "in.bgr"
    raw.write_bytes(frames.tobytes())
    r = subprocess.run([ENC, "--width", str(w), "--height", str(h), "--batch", "3", "--verbose", "0",
                        "--seed", "3", "--transform-block-w", str(tb), "--transform-block-h", str(tb), str(raw)],
                       capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr
    stream = r.stdout
    fg, bg = 2, 24
    base = ["--foreground-quant-step", str(fg), "--background-quant-step", str(bg), "--batch", "3",
            "--verbose", "0"]
    if gaze:
        base += ["--gaze-x", str(gaze[0]), "--gaze-y", str(gaze[1])]
    d = subprocess.run([DEC] + base + ["--padded", "1", "-"], input=stream, capture_output=True, timeout=300)
    assert d.returncode == 0, d.stderr
    pw, ph = oracle.padded_dim(w, 16, 4), oracle.padded_dim(h, 16, 4)
    out = np.frombuffer(d.stdout, np.uint8).reshape(n - 1, ph, pw, 3)
    fb = oracle.serialized_frame_bytes(w, h, tb, tb)
    rows = (h + tb - 1) // tb * tb  # block rows the encoder wrote
    gz = oracle.gaze_rect(gaze[0], gaze[1], 64, 64, w, h, pw, ph) if gaze else None
    if gz is not None and gz[1] + gz[3] > rows:  # the oracle decodes the written rows only
        gz = (gz[0], gz[1], gz[2], max(0, rows - gz[1]))
    step = 1 if w * h < 500 * 500 else n - 2   # the oracle's per-block loop is slow at 1080p
    for t in range(0, n - 1, step):
        rec = np.frombuffer(stream, np.uint8, fb, 32 + t * fb)
        exp = oracle.decode_frame_blocks(rec, pw, rows, tb, tb, fg_q=fg, bg_q=bg, gaze=gz)
        exp8 = np.clip(np.rint(exp), 0, 255)
        got = out[t, :rows].astype(np.float32)
        # a float error <= 1e-3 can flip the 8-bit rounding only next to .5: allow those, nothing else
        diff = np.abs(got - exp8)
        near_half = np.abs(np.abs(exp - np.floor(exp)) - 0.5) < 2e-3
        assert diff.max() <= 1 and not (diff > 0)[~near_half].any()
        assert not out[t, rows:].any()
    # cropped output (the default): the same pixels without the padding
    d2 = subprocess.run([DEC] + base + ["-"], input=stream, capture_output=True, timeout=300)
    assert d2.returncode == 0, d2.stderr
    crop = np.frombuffer(d2.stdout, np.uint8).reshape(n - 1, h, w, 3)
    assert np.array_equal(crop, out[:, :h, :w])
