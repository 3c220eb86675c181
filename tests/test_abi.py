"""CPU: the C-ABI library loads, exports every symbol include/svc_b200.h
declares, validates arguments like the reference, and fails loudly (never
falls back) where no CUDA device exists.  No compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "svc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svc_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(svc):
    L = svc.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/svc_b200.h but not exported"


def test_version_and_geometry(svc):
    assert b"sm_100a" in svc.lib().svc_version()
    assert svc.padded_dim(1080, 16, 4) == 1088
    assert svc.padded_dim(960, 16, 4) == 960
    assert svc.serialized_frame_bytes(1920, 1080) == 240 * 135 * 772
    assert svc.serialized_frame_bytes(960, 540) == 120 * 68 * 772
    hdr = svc.write_header(30, 960, 540, 960, 544).view(np.uint32)
    assert list(hdr) == [29, 960, 540, 0, 4, 8, 8, 3]


def test_header_and_sizes_agree_with_oracle(svc, oracle):
    for (n, w, h) in ((300, 1920, 1080), (2, 104, 56), (1, 16, 16), (0, 8, 8)):
        pw, ph = svc.padded_dim(w, 16, 4), svc.padded_dim(h, 16, 4)
        assert (pw, ph) == (oracle.padded_dim(w, 16, 4), oracle.padded_dim(h, 16, 4))
        assert np.array_equal(svc.write_header(n, w, h, pw, ph), oracle.header(n, w, h, pw, ph))
        assert svc.serialized_frame_bytes(w, h) == oracle.serialized_frame_bytes(w, h)


def _pyr(w, h, L):
    return [np.zeros((h >> l, w >> l), np.uint8) for l in range(L)]


@pytest.mark.parametrize("kwargs,msg", [
    (dict(level_count=0), "level_count"),
    (dict(level_count=9), "level_count"),
    (dict(block_w=0), "> 0"),
    (dict(frame_w=100), "divisible by the block"),
    (dict(search_range=4), "search_range must be >="),
    (dict(block_w=12, frame_w=96), "divisible by 2^"),   # Q15: reference would divide by zero
])
def test_hbma_preconditions(svc, kwargs, msg):
    a = dict(level_count=4, frame_w=64, frame_h=64, search_range=8, block_w=16, block_h=16)
    a.update(kwargs)
    L = max(1, min(a["level_count"], 9))
    p = _pyr(128, 128, L)
    with pytest.raises(svc.SvcError) as e:
        svc.EstimateMotionHierarchical(p, p, **a)
    assert e.value.code == 1 and msg in str(e.value)


def test_session_config_validation_mirrors_reference_messages(svc):
    bad = [
        (dict(mv_block_w=0), "invalid mv block width"),
        (dict(pyr_lvl_count=0), "invalid pyramid level count"),
        (dict(mv_search_range=7), "quotient"),
        (dict(transform_block_w=32), "transform block width must be <= mv block width"),
        (dict(transform_block_h=5), "mv block height must be divisible"),
        (dict(frame_w=0), "frame dimensions"),
    ]
    for kw, msg in bad:
        cfg = svc.SessionConfig(frame_w=64, frame_h=64)
        for k, v in kw.items():
            setattr(cfg, k, v)
        with pytest.raises(svc.SvcError) as e:
            svc.Session(cfg)
        assert e.value.code == 1 and msg in str(e.value), (kw, str(e.value))


def test_session_config_struct_size_versions_the_struct(svc):
    """struct_size versions svc_session_config: a caller compiled against the round-1 header (no
    hbma_kernel_family / host_chunk_frames) is accepted with those fields zero; sizes that match no
    version and unknown kernel families are argument errors."""
    from svc_b200.binding import _Cfg
    L = svc.lib()
    old_size = _Cfg.hbma_kernel_family.offset

    def create(struct_size, family=0, w=64):
        c = _Cfg(struct_size, w, 64, 16, 16, 8, 4, 8, 8, 0, 0, None, family, 0)
        h = C.c_void_p()
        rc = L.svc_session_create(C.byref(c), C.byref(h))
        msg = L.svc_last_error().decode()
        if rc == 0:
            L.svc_session_destroy(h)
        return rc, msg

    for size in (old_size, C.sizeof(_Cfg)):
        rc, msg = create(size, w=0)  # passes the size check, then fails on the frame width
        assert rc == 1 and "frame dimensions" in msg, (size, msg)
    for size in (old_size - 4, C.sizeof(_Cfg) + 8, 0):
        rc, msg = create(size)
        assert rc == 1 and "struct_size" in msg, (size, msg)
    rc, msg = create(C.sizeof(_Cfg), family=9)
    assert rc == 1 and "hbma_kernel_family" in msg
    rc, msg = create(old_size, family=9)  # the field lies beyond the declared size: ignored
    assert "hbma_kernel_family" not in msg


def test_no_device_is_a_hard_error_not_a_fallback(svc):
    if svc.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    p = _pyr(64, 64, 4)
    with pytest.raises(svc.SvcError) as e:
        svc.EstimateMotionHierarchical16x16Sse2(p, p, 64, 64, 8)
    assert e.value.code == 2
    with pytest.raises(svc.SvcError) as e:
        svc.Session(svc.SessionConfig(frame_w=64, frame_h=64))
    assert e.value.code == 2
    with pytest.raises(svc.SvcError):
        svc.dct_planar(np.zeros((16, 16, 3), np.uint8), 16, 16)


def test_patch_block_types_matches_oracle_serializer(svc, oracle):
    rng = np.random.default_rng(5)
    w, h, pw, ph = 96, 40, 96, 48
    planes = rng.random((3, ph, pw), dtype=np.float32)
    bt = rng.integers(0, 7, size=(ph // 16) * (pw // 16)).astype(np.uint32)
    ref_zero = oracle.serialize_frame(planes, None, w, h, 8, 8, pw // 16, 16, 16)
    ref_bt = oracle.serialize_frame(planes, bt, w, h, 8, 8, pw // 16, 16, 16)
    st = ref_zero.copy()
    svc.patch_block_types(st, w, h, bt, mv_field_w=pw // 16)
    assert np.array_equal(st, ref_bt)


def test_stream_layout_reports_the_reference_encoder_decoder_mismatch(svc):
    # encoder iterates the unpadded frame, the reference decoder the padded one
    l = svc.stream_layout(svc.write_header(300, 1920, 1080, 1920, 1088))
    assert (l["encoder_records_per_frame"], l["decoder_records_per_frame"]) == (240 * 135, 240 * 136)
    assert l["consistent"] == 0 and l["record_bytes"] == 772
    assert l["encoder_stream_bytes"] == 32 + 299 * 240 * 135 * 772
    for (w, h) in ((960, 540), (3840, 2160)):
        pw, ph = svc.padded_dim(w, 16, 4), svc.padded_dim(h, 16, 4)
        assert svc.stream_layout(svc.write_header(30, w, h, pw, ph))["consistent"] == 1
    with pytest.raises(svc.SvcError):
        svc.stream_layout(np.zeros(32, np.uint8))


def test_header_constants_match_the_binding(svc):
    """Every SVC_HBMA_FAMILY_* / SVC_STAGE_* value of include/svc_b200.h is the value the ctypes binding
    uses (the test hook that routes the default configuration to the tile or the strip kernels
    depends on it), and the session rejects a family beyond the last one defined."""
    src = open(os.path.join(ROOT, "include", "svc_b200.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(SVC_(?:HBMA_FAMILY|STAGE)_[A-Z0-9_]+)\s+(\d+)u?", src)}
    fam = {k[len("SVC_"):]: v for k, v in defs.items() if k.startswith("SVC_HBMA_FAMILY_")}
    assert set(fam) >= {"HBMA_FAMILY_AUTO", "HBMA_FAMILY_GENERIC", "HBMA_FAMILY_POOL", "HBMA_FAMILY_WINDOW", "HBMA_FAMILY_TILE"}
    for name, value in fam.items():
        assert getattr(svc, name) == value, name
    assert sorted(fam.values()) == list(range(len(fam)))
    for name, value in defs.items():
        if name.startswith("SVC_STAGE_"):
            assert getattr(svc, name[len("SVC_"):]) == value, name
