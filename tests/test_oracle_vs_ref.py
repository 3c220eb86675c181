"""CPU: the C oracle against the compiled, unmodified reference motion code on
fresh seeded inputs (skipped where oracle/_ref was never built)."""
import numpy as np
import pytest

from svc_b200.synth import SyntheticSequence


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libref_motion.so not built (needs /root/reference)")
    return oracle


def _pyr(oracle, f, pw, ph, L):
    return oracle.y_pyramid(f, pw, ph, L)


@pytest.mark.parametrize("w,h,seed", [(320, 180, 1), (208, 112, 2), (64, 48, 3), (16, 16, 4)])
@pytest.mark.parametrize("R", [8, 17, 40])
def test_sse2_entry(ref, w, h, seed, R):
    pw, ph = ref.padded_dim(w, 16, 4), ref.padded_dim(h, 16, 4)
    seq = SyntheticSequence(w, h, 2, seed=seed, n_rects=3)
    p0, p1 = _pyr(ref, seq.frame(0), pw, ph, 4), _pyr(ref, seq.frame(1), pw, ph, 4)
    mo, ao = ref.hbma(p0, p1, R)
    for impl in ("ref", "ref_sse2"):
        mr, ar = ref.hbma(p0, p1, R, impl=impl)
        assert np.array_equal(mo, mr) and np.array_equal(ao, ar)


@pytest.mark.parametrize("L,bw,bh,R", [(1, 16, 16, 3), (2, 8, 8, 9), (3, 32, 16, 12),
                                       (5, 16, 16, 16), (2, 6, 10, 7), (1, 3, 7, 1)])
def test_generic_entry(ref, L, bw, bh, R):
    rng = np.random.default_rng(L * 100 + bw)
    w, h = bw * 11, bh * 6
    f = SyntheticSequence(w, h, 2, seed=bw * 3 + bh, n_rects=2)
    pw, ph = ref.padded_dim(w, bw, L), ref.padded_dim(h, bh, L)
    p0, p1 = _pyr(ref, f.frame(0), pw, ph, L), _pyr(ref, f.frame(1), pw, ph, L)
    # sprinkle exact ties
    p1[0][: bh, :] = p0[0][: bh, :] = rng.integers(0, 2) * 200
    mo, ao = ref.hbma(p0, p1, R, bw, bh)
    mr, ar = ref.hbma(p0, p1, R, bw, bh, impl="ref")
    assert np.array_equal(mo, mr) and np.array_equal(ao, ar)


def test_flat_and_zero_frames(ref):
    z = [np.zeros((64 >> l, 96 >> l), np.uint8) for l in range(4)]
    c = [np.full((64 >> l, 96 >> l), 9, np.uint8) for l in range(4)]
    for t, a in ((z, z), (z, c), (c, z)):
        mo, ao = ref.hbma(t, a, 8)
        mr, ar = ref.hbma(t, a, 8, impl="ref_sse2")
        assert np.array_equal(mo, mr) and np.array_equal(ao, ar)
        assert not mo.any()  # non-increasing sequence -> zero vector at the top, no strict gain below
