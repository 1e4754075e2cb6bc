"""The arithmetic the tensor-core paths rest on, checked on the CPU with numpy (no device):
an fp32 value is carried as bf16 hi + bf16 lo (hi = rn(x), lo = rn(x - hi)); the forward kernel sums hi*hi + hi*lo +
lo*hi (spconv_v6.cu), the filter gradient all four products (spconv_wgrad_tc.cu), a residual is read back as hi + lo
(encoder.cu).  These bounds are what lets both stay inside the 1e-4 parity bar with fp32 accumulation."""
import numpy as np


def bf16_rn(x):
    """round-to-nearest-even to bfloat16, returned as float32 (what cvt.rn.bf16x2.f32 does for finite values)"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split(x):
    hi = bf16_rn(x)
    lo = bf16_rn(np.asarray(x, np.float32) - hi)
    return hi, lo


def test_hi_plus_lo_reproduces_the_value_to_2_pow_minus_16():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-20, 20, 200000))).astype(np.float32)
    hi, lo = split(x)
    # the fp32 sum hi + lo is exact (8 + 8 significant bits, adjacent exponents) and within 2^-16 of x
    s = hi.astype(np.float64) + lo.astype(np.float64)
    assert np.array_equal((hi + lo).astype(np.float64), s)
    assert np.max(np.abs(s - x) / np.abs(x)) <= 2.0 ** -16


def test_three_and_four_product_forms_against_the_fp32_product():
    rng = np.random.default_rng(1)
    a = rng.standard_normal(200000).astype(np.float32)
    w = rng.standard_normal(200000).astype(np.float32)
    (ah, al), (wh, wl) = split(a), split(w)
    exact = a.astype(np.float64) * w.astype(np.float64)
    f64 = lambda v: v.astype(np.float64)
    three = f64(ah) * f64(wh) + f64(ah) * f64(wl) + f64(al) * f64(wh)
    four = three + f64(al) * f64(wl)
    scale = np.abs(exact)
    # dropping lo*lo costs at most 2^-16 (|lo| <= 2^-9 |x| on both sides); with it only the two splits' own error is left
    assert np.max(np.abs(three - exact) / scale) <= 2.0 ** -14
    assert np.max(np.abs(four - exact) / scale) <= 2.0 ** -15
    # and the sums the kernels actually form: a K = 27 * 128 dot product stays far inside the 1e-4 bar
    k = 27 * 128
    A = rng.standard_normal((64, k)).astype(np.float32)
    W = rng.standard_normal((k, 32)).astype(np.float32)
    (Ah, Al), (Wh, Wl) = split(A), split(W)
    gold = f64(A) @ f64(W)
    got = (f64(Ah) @ f64(Wh) + f64(Ah) @ f64(Wl) + f64(Al) @ f64(Wh)).astype(np.float32)
    assert np.max(np.abs(got - gold)) <= 2e-5 * np.max(np.abs(gold))
