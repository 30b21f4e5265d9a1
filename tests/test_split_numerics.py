"""Host-side restatement (numpy) of the operand split the tensor-core data passes use (gppvae_b200/csrc/pass1_common.cuh
hi11_round / kF16Top, gemm_planes.cu split4; the converters of gemm_tc.cu use the same arithmetic): checks the error
bounds DESIGN.md 5.1 states for it, independent of any GPU.  A specification test of the arithmetic, not of the kernels (those are covered by the
`-m gpu` parity tests)."""
import numpy as np
import pytest

K_F16_TOP = 7   # gemm_tc.cu kF16Top


def exp_of_max(x):
    """e with 2^(e-1) <= max|x| < 2^e (gemm_tc.cu exp_of_bits on the absmax bit pattern)."""
    m = np.float32(np.abs(x).max())
    if m == 0:
        return 0
    bits = int(np.array(m, dtype=np.float32).view(np.uint32))
    return ((bits >> 23) & 0xFF) - 126


def hi11(x):
    """x rounded to 11 significant bits, half away from zero: (bits + 0x1000) & 0xFFFFE000."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x, e):
    """(hi, lo) as the fp16 numbers the converters write, and the power-of-two scale they carry."""
    x = np.asarray(x, dtype=np.float32)
    s = np.float32(2.0 ** (K_F16_TOP - e))
    h = hi11(x)
    lo = (x - h).astype(np.float32)          # exact in fp32
    with np.errstate(over="ignore"):
        h16 = (h * s).astype(np.float16)     # cvt.rn.f16.f32 (satfinite on the device; inf here marks saturation)
        l16 = (lo * s).astype(np.float16)
    return h16, l16, float(s)


def three_term_product(a, b):
    ea, eb = exp_of_max(a), exp_of_max(b)
    ah, al, sa = split(a, ea)
    bh, bl, sb = split(b, eb)
    f = lambda v: v.astype(np.float64)
    return (f(ah) * f(bh) + f(ah) * f(bl) + f(al) * f(bh)) / (sa * sb)


@pytest.mark.parametrize("scale_a,scale_b", [(1.0, 1.0), (1e-6, 1e-6), (3e5, 2e4), (1e-4, 1e3)])
def test_three_terms_reach_fp32_accuracy_whatever_the_units(scale_a, scale_b):
    rng = np.random.default_rng(0)
    a = (rng.standard_normal(200_000) * scale_a).astype(np.float32)
    b = (rng.standard_normal(200_000) * scale_b).astype(np.float32)
    got = three_term_product(a, b)
    ref = a.astype(np.float64) * b.astype(np.float64)
    big = (np.abs(a) > 2.0 ** -10 * np.abs(a).max()) & (np.abs(b) > 2.0 ** -10 * np.abs(b).max())
    rel = np.abs(got - ref)[big] / np.abs(ref)[big]
    # dropped lo.lo <= 2^-22, two lo roundings <= 2^-23 each (+ the hi factor's exactness): well below 2^-20
    assert rel.max() < 2.0 ** -20
    assert rel.mean() < 2.0 ** -23
    # small elements: absolute precision relative to the operand maxima (DESIGN: 2^-32 of the maximum per operand)
    err = np.abs(got - ref)[~big]
    bound = 2.0 ** -30 * float(np.abs(a).max()) * float(np.abs(b).max())
    assert err.size == 0 or err.max() < bound


def test_hi_is_exactly_an_fp16_number_and_lo_is_zero_mean():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(100_000).astype(np.float32)
    e = exp_of_max(x)
    h16, l16, s = split(x, e)
    big = np.abs(x) > 2.0 ** -10 * np.abs(x).max()
    assert np.array_equal(h16.astype(np.float32)[big], (hi11(x) * np.float32(s))[big])      # no second rounding
    lo = (x - hi11(x)).astype(np.float64)
    assert np.all(np.abs(lo) <= 2.0 ** -11 * np.abs(x) * (1 + 1e-6))                        # rounded, not truncated
    assert abs((lo / x)[big].mean()) < 2.0 ** -11 / 50                                      # zero-mean remainder


def test_headroom_above_the_maximum():
    """The scale is derived from the exact maximum, so nothing can exceed it; the format itself still leaves 2^8 of
    headroom (an element up to 2^8 above the maximum the scale was derived from converts without saturating)."""
    x = np.array([1.0, 0.3, -0.9], dtype=np.float32)
    e = exp_of_max(x)                                   # 2^0 <= 1.0 < 2^1  ->  e = 1
    outlier = np.array([250.0 * 2.0 ** (e - 1)], dtype=np.float32)
    h16, l16, _ = split(outlier, e)
    assert np.isfinite(h16).all() and np.isfinite(l16).all()
    too_big = np.array([2.0 ** (e + 9)], dtype=np.float32)
    h16, _, _ = split(too_big, e)
    assert not np.isfinite(h16).all()                   # (the device saturates to 65504 instead of inf)
