"""Block-type stages that consume the motion field (reference
libs/encoder.cpp:491-624, libs/motion.cpp:157-266; product host/segment.cpp behind
include/svc_segment.h).  CPU only.

Checkers: the UNMODIFIED reference EstimateGlobalMotionRansac compiled into oracle/_ref (its
function-static engine seeded through an interposed std::random_device), python cv2 for the
OpenCV calls (morphologyEx, kmeans, connectedComponents), and the committed fixtures of
tests/golden/make_golden_segment.py generated from the same two sources."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

needs_cv2 = pytest.mark.skipif(cv2 is None, reason="python cv2 not importable")


@pytest.fixture(scope="module")
def seg(svc):  # svc: builds the libraries on a fresh checkout
    from svc_b200 import segment
    segment.host_lib()
    return segment


@pytest.fixture(scope="module")
def gold():
    return load_golden("segment.npz")


def _need_ref(oracle, seed):
    L = oracle.ref_seeded(seed)
    if L is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return L


def synth_mv_field(w, h, seed):
    rng = np.random.default_rng(seed)
    mv = np.zeros((h, w, 2), np.float32)
    mv[:] = rng.integers(-3, 4, size=2)
    for _ in range(int(rng.integers(2, 6))):
        x0, y0 = int(rng.integers(0, w - 4)), int(rng.integers(0, h - 4))
        x1, y1 = min(w, x0 + int(rng.integers(3, max(4, w // 3)))), min(h, y0 + int(rng.integers(3, max(4, h // 3))))
        mv[y0:y1, x0:x1] = rng.integers(-24, 25, size=2)
    noise = rng.random((h, w)) < 0.04
    mv[noise] += rng.integers(-16, 17, size=(int(noise.sum()), 2))
    return mv


# ------------------------------------------------------------------ boundary
def test_host_library_exports_every_declared_symbol(seg):
    src = open(os.path.join(ROOT, "include", "svc_segment.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(svc_seg_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 11
    for n in names:
        assert hasattr(seg.host_lib(), n), f"{n} declared in include/svc_segment.h but not exported"


@pytest.mark.parametrize("field,value,msg", [
    ("ransac_inlier_thresh", -1.0, "invalid inlier threshold: must be >= 0"),
    ("ransac_success_prob", -0.5, "invalid success probability: must be >= 0"),
    ("ransac_inlier_ratio", -0.1, "invalid inlier ratio: must be >= 0"),
    ("kmeans_cluster_count", 0, "invalid cluster count: must be > 0"),
    ("kmeans_attempt_count", 0, "invalid attempt count: must be > 0"),
    ("kmeans_max_iter_count", 0, "invalid maximum iteration count: must be > 0"),
    ("kmeans_epsilon", 0.0, "invalid epsilon: must be > 0"),
    ("connected_components_connectivity", 6, "invalid connected components connectivity: must be either 4 or 8"),
])
def test_validate_messages_match_the_reference(seg, field, value, msg):
    """libs/encoder.cpp:20-60, 96-101"""
    cfg = seg.SegmentConfig()
    assert seg.validate(cfg) == ""
    setattr(cfg, field, value)
    assert seg.validate(cfg) == msg
    with pytest.raises(seg.SegError):
        seg.block_types(np.zeros((4, 4, 2), np.float32), cfg)


# ------------------------------------------------------------------ RANSAC
def test_ransac_golden(seg, gold):
    for ci in range(3):
        state = int(gold[f"ransac{ci}_0_seed"][0])
        for call in range(2):
            p = gold[f"ransac{ci}_{call}_params"]
            rmse, gm, inl, state = seg.ransac(gold[f"ransac{ci}_{call}_mv"], int(p[0]), p[1], p[2], p[3], state)
            assert np.array_equal(inl, gold[f"ransac{ci}_{call}_inliers"])
            assert np.array_equal(gm, gold[f"ransac{ci}_{call}_gm"])
            assert np.float32(rmse) == gold[f"ransac{ci}_{call}_rmse"][0]


@pytest.mark.parametrize("seed", [1, 2, 424242, 2 ** 31 - 2, 0])
@pytest.mark.parametrize("params", [(1, 7.5, 0.99, 0.5), (4, 3.0, 0.999, 0.45), (2, 0.5, 0.5, 0.9)])
def test_ransac_matches_the_compiled_reference(seg, oracle, seed, params):
    """Same seed -> same subsets, inliers, refit and rmse, call after call (the engine state is
    carried over like the reference's function-static engine)."""
    L = _need_ref(oracle, seed)
    state = seed
    for call in range(3):
        mv = synth_mv_field(60, 34, seed % 1000 + call)
        e_rmse, e_gm, e_inl = oracle.ref_ransac(L, mv, *params, gm0=(0.5, -0.25))
        rmse, gm, inl, state = seg.ransac(mv, *params, rng_state=state, gm0=(0.5, -0.25))
        if np.isnan(e_gm).any() or np.isnan(e_rmse):
            # the reference sampled index n (one past the field, libs/motion.cpp:208; the checker
            # fills that slot with NaN) in its winning iteration: undefined there, dropped here --
            # the engine state still has to stay in step for the next call
            continue
        assert np.array_equal(inl, e_inl) and np.array_equal(gm, e_gm)
        assert np.float32(rmse) == np.float32(e_rmse)


def test_ransac_draw_one_past_the_field_is_dropped(seg, oracle):
    """Pins the one deliberate deviation of the RANSAC restatement: the reference draws subset indices
    from [0, n] INCLUSIVE (libs/motion.cpp:208) and, when n comes up, reads one element past the motion
    field -- undefined behaviour.  Here that iteration is dropped after consuming the same draws.  With
    5-vector fields a draw of n happens in most calls (7 iterations x 1/6 each).  The checker stores a
    far-away vector in the slot the reference over-reads, so such an iteration has exactly one inlier-free
    subset mean and can only win while nothing else has been evaluated; every call whose winning subset
    is inside the field must then agree bit for bit, call after call (the engine state stays in step)."""
    import ctypes as C
    L = _need_ref(oracle, 99)
    state, hit, agree = 99, 0, 0
    for call in range(60):
        rng = np.random.default_rng(call)
        mv = rng.integers(-3, 4, size=(5, 2)).astype(np.float32)
        n = mv.shape[0]
        buf = np.zeros((n + 1, 2), np.float32)
        buf[:n] = mv
        buf[n] = 1.0e6  # what the reference reads when it draws n: no vector is within 7.5 of it
        rm, ni = C.c_float(), C.c_uint32()
        gm = np.zeros(2, np.float32)
        inl = np.zeros(n, np.uint32)
        f32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        L.ref_ransac(buf.ctypes.data_as(f32p), n, 1, C.c_float(7.5), C.c_float(0.99), C.c_float(0.5), C.byref(rm),
                     gm.ctypes.data_as(f32p), inl.ctypes.data_as(u32p), C.byref(ni))
        rmse, g, i2, state = seg.ransac(mv, 1, 7.5, 0.99, 0.5, rng_state=state)
        if abs(gm).max() > 1e5:  # the over-read slot won in the reference (every earlier draw was n too)
            hit += 1
            continue
        agree += 1
        assert np.array_equal(i2, inl[:ni.value]) and np.array_equal(g, gm) and np.float32(rmse) == np.float32(rm.value)
    assert agree >= 50


def test_ransac_degenerate_no_consensus(seg, oracle):
    """inlier_thresh 0: no vector is ever an inlier; the reference then reports the rmse of the last
    subset against the CALLER's global_motion value (libs/motion.cpp:239-241)."""
    L = _need_ref(oracle, 5)
    mv = synth_mv_field(20, 12, 3)
    e = oracle.ref_ransac(L, mv, 2, 0.0, 0.99, 0.5, gm0=(1.0, 2.0))
    g = seg.ransac(mv, 2, 0.0, 0.99, 0.5, rng_state=5, gm0=(1.0, 2.0))
    assert e[2].size == 0 and g[2].size == 0
    assert np.array_equal(g[1], e[1]) and np.float32(g[0]) == np.float32(e[0])


def test_global_motion_avg_matches_the_reference(seg, oracle):
    import ctypes as C
    L = _need_ref(oracle, 1)
    for seed in range(4):
        mv = synth_mv_field(33, 21, seed).reshape(-1, 2)
        e = np.zeros(2, np.float32)
        L.ref_global_motion_avg(mv.ctypes.data_as(C.POINTER(C.c_float)), mv.shape[0], e.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(seg.global_motion_avg(mv), e)


def test_ransac_rejects_small_fields(seg):
    with pytest.raises(seg.SegError):
        seg.ransac(np.zeros((2, 2), np.float32), subset_sz=3)


# ------------------------------------------------------------------ morphology / connected components
def test_morphology_and_components_golden(seg, gold):
    for i in range(4):
        m = gold[f"morph{i}_mask"]
        rw, rh = (int(v) for v in gold[f"morph{i}_rect"])
        for op in range(4):
            assert np.array_equal(seg.morphology(m, op, rw, rh), gold[f"morph{i}_op{op}"])
        for conn in (4, 8):
            n, lab = seg.connected_components(m, conn)
            assert n == int(gold[f"cc{i}_conn{conn}_n"][0]) and np.array_equal(lab, gold[f"cc{i}_conn{conn}_labels"])


@needs_cv2
@pytest.mark.parametrize("w,h", [(120, 68), (7, 5), (1, 9), (33, 2), (240, 135)])
def test_morphology_matches_cv2(seg, w, h):
    rng = np.random.default_rng(w * 131 + h)
    for (rw, rh) in ((3, 3), (5, 3), (2, 2), (4, 3), (1, 1), (1, 3), (7, 7)):
        el = cv2.getStructuringElement(cv2.MORPH_RECT, (rw, rh))
        for op in (seg.MORPH_ERODE, seg.MORPH_DILATE, seg.MORPH_OPEN, seg.MORPH_CLOSE):
            for dens in (0.1, 0.5, 0.9):
                m = ((rng.random((h, w)) < dens) * 255).astype(np.uint8)
                assert np.array_equal(seg.morphology(m, op, rw, rh), cv2.morphologyEx(m, op, el))
    for m in (np.zeros((h, w), np.uint8), np.full((h, w), 255, np.uint8)):
        assert np.array_equal(seg.morphology(m, seg.MORPH_CLOSE), m)


@needs_cv2
@pytest.mark.parametrize("conn", [4, 8])
@pytest.mark.parametrize("w,h", [(120, 68), (7, 5), (1, 9), (33, 2), (61, 35), (240, 135)])
def test_connected_components_match_cv2(seg, conn, w, h):
    """Label VALUES (OpenCV's numbering order), not just the partition: the block types written to
    the stream are label + offset (libs/encoder.cpp:613-621)."""
    rng = np.random.default_rng(conn * 1000 + w + h)
    masks = [((rng.random((h, w)) < d) * 255).astype(np.uint8) for d in (0.05, 0.3, 0.5, 0.6, 0.8) for _ in range(4)]
    masks += [np.zeros((h, w), np.uint8), np.full((h, w), 255, np.uint8)]
    chk = np.zeros((h, w), np.uint8)
    chk[::2, ::2] = 255
    chk[1::2, 1::2] = 255
    masks.append(chk)  # checkerboard: one component (8) / all singletons (4)
    for m in masks:
        e_n, e_lab = cv2.connectedComponents(m, connectivity=conn, ltype=cv2.CV_32S)
        n, lab = seg.connected_components(m, conn)
        assert n == e_n and np.array_equal(lab, e_lab)


# ------------------------------------------------------------------ k-means
def test_kmeans_golden(seg, gold):
    for i in range(4):
        k, max_iter, eps, attempts, seed = gold[f"kmeans{i}_args"]
        comp, lab, cen, _ = seg.kmeans(gold[f"kmeans{i}_data"], int(k), int(max_iter), float(eps), int(attempts), int(seed))
        assert np.array_equal(lab, gold[f"kmeans{i}_labels"])
        assert np.allclose(cen, gold[f"kmeans{i}_centers"], rtol=0, atol=1e-4)
        assert abs(comp - gold[f"kmeans{i}_compactness"][0]) <= 1e-6 * max(1.0, abs(comp))


@needs_cv2
def test_kmeans_matches_cv2_labels(seg):
    """cv::kmeans with KMEANS_PP_CENTERS from the same cv::RNG state: identical labels (so identical
    centre seeding, empty-cluster handling, termination and best-attempt choice)."""
    rng = np.random.default_rng(2)
    for trial in range(120):
        n = int(rng.integers(1, 1200))
        k = int(min(n, rng.integers(1, 12)))
        kind = trial % 3
        if kind == 0:
            data = (rng.normal(size=(n, 4)) * 50).astype(np.float32)
        elif kind == 1:  # the encoder's features: (0, mv.x, block x, block y)
            idx = np.sort(rng.choice(8160, size=n, replace=False))
            data = np.zeros((n, 4), np.float32)
            data[:, 1] = rng.integers(-8, 9, size=n)
            data[:, 2] = (idx % 120) * 16
            data[:, 3] = (idx // 120) * 16
        else:  # duplicates and tight blobs: empty clusters do happen
            c = rng.integers(-50, 50, size=(3, 4))
            data = c[rng.integers(0, 3, size=n)].astype(np.float32)
        max_iter = int(rng.integers(1, 15))
        eps = float(rng.choice([1.0, 0.1, 5.0]))
        attempts = int(rng.integers(1, 4))
        seed = int(rng.integers(1, 2 ** 31))
        cv2.setRNGSeed(seed)
        e_comp, e_lab, e_cen = cv2.kmeans(data.reshape(n, 1, 4), k, None,
                                          (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, max_iter, eps), attempts,
                                          cv2.KMEANS_PP_CENTERS)
        comp, lab, cen, _ = seg.kmeans(data, k, max_iter, eps, attempts, seed)
        assert np.array_equal(lab, e_lab.reshape(-1)), (trial, n, k)
        assert np.allclose(cen, e_cen, rtol=0, atol=1e-3)
        assert abs(comp - e_comp) <= 1e-6 * max(1.0, abs(e_comp))


@needs_cv2
def test_kmeans_state_carries_over_like_theRNG(seg):
    """Two consecutive cv2.kmeans calls share cv::theRNG(); the returned state reproduces that."""
    rng = np.random.default_rng(8)
    a = (rng.normal(size=(300, 4)) * 20).astype(np.float32)
    b = (rng.normal(size=(200, 4)) * 20).astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 10, 1.0)
    cv2.setRNGSeed(77)
    _, la, _ = cv2.kmeans(a.reshape(-1, 1, 4), 6, None, crit, 3, cv2.KMEANS_PP_CENTERS)
    _, lb, _ = cv2.kmeans(b.reshape(-1, 1, 4), 5, None, crit, 3, cv2.KMEANS_PP_CENTERS)
    _, ga, _, st = seg.kmeans(a, 6, 10, 1.0, 3, 77)
    _, gb, _, _ = seg.kmeans(b, 5, 10, 1.0, 3, st)
    assert np.array_equal(ga, la.reshape(-1)) and np.array_equal(gb, lb.reshape(-1))


# ------------------------------------------------------------------ the whole chain
def test_block_types_golden(seg, gold):
    for i in range(3):
        rseed, kseed, conn = (int(v) for v in gold[f"chain{i}_seeds"])
        cfg = seg.SegmentConfig(connected_components_connectivity=conn)
        bt, gm, _, _ = seg.block_types(gold[f"chain{i}_mv"], cfg, rseed, kseed)
        assert np.array_equal(bt, gold[f"chain{i}_types"])
        assert np.array_equal(gm, gold[f"chain{i}_gm"])
        assert bt.max() > 0


@needs_cv2
@pytest.mark.parametrize("w,h,conn", [(120, 68, 4), (120, 68, 8), (60, 34, 4), (20, 12, 8), (240, 135, 4)])
def test_block_types_match_reference_ransac_plus_cv2(seg, oracle, w, h, conn):
    """libs/encoder.cpp:491-624 end to end: inliers from the compiled reference, every OpenCV stage
    from cv2, against svc_seg_block_types from the same generator states."""
    for rep in range(3):
        rseed, kseed = 1000 + 7 * rep + w, 50 + rep
        L = _need_ref(oracle, rseed)
        mv = synth_mv_field(w, h, w * 3 + rep)
        _, e_gm, inl = oracle.ref_ransac(L, mv)
        cfg = seg.SegmentConfig(connected_components_connectivity=conn, kmeans_cluster_count=10 - 3 * rep)
        exp = oracle.block_types_cv2(mv, inl, kseed, connectivity=conn, cluster_count=10 - 3 * rep)
        bt, gm, _, _ = seg.block_types(mv, cfg, rseed, kseed)
        assert np.array_equal(bt, exp) and np.array_equal(gm, e_gm)


def test_block_types_all_background_when_motion_is_global(seg):
    mv = np.zeros((34, 60, 2), np.float32)
    mv[:] = (3, -2)
    bt, gm, _, _ = seg.block_types(mv)
    assert not bt.any() and np.allclose(gm, (3.0, -2.0), atol=1e-5)  # sum * (1.0f / n), as the reference


def test_batch_stage_is_independent_of_threads_and_batching(seg):
    """svc::BlockTypeStage: per-frame generators derived from (seed, frame index) -> the labels do
    not depend on worker threads, batch boundaries or the first frame of a shard."""
    fields = np.stack([synth_mv_field(60, 34, 40 + i) for i in range(9)])
    one = seg.block_types_batch(fields, seed=12345, first_frame=0, threads=1)
    many = seg.block_types_batch(fields, seed=12345, first_frame=0, threads=8)
    assert np.array_equal(one, many) and one.any()
    split = np.concatenate([seg.block_types_batch(fields[:4], seed=12345, first_frame=0, threads=3),
                            seg.block_types_batch(fields[4:], seed=12345, first_frame=4, threads=2)])
    assert np.array_equal(one, split)
    for f in range(9):  # and it is exactly the single-field entry point with the published generator states
        r, k = seg.frame_generators(12345, f)
        assert np.array_equal(seg.block_types(fields[f], None, r, k)[0], one[f])
    assert not np.array_equal(one, seg.block_types_batch(fields, seed=54321, first_frame=0, threads=2))
