"""The fixed-point accumulation behind the deterministic-statistics mode (csrc/common.cuh: det_add / det_value),
compiled for the host from the same source (tests/emu_harness.py): the sum of a set of partials must not depend on
the order they are added in, and must equal the exact sum to the stated resolution.  CPU only."""
import ctypes as C
from fractions import Fraction

import numpy as np
import pytest

from emu_harness import load_emu


def _det_sum(lib, v, order):
    v = np.ascontiguousarray(v, dtype=np.float64)
    order = np.ascontiguousarray(order, dtype=np.int32)
    d, f = C.c_double(), C.c_float()
    lib.tdanet_emu_det_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    assert lib.tdanet_emu_det_sum(v.ctypes.data, order.ctypes.data, len(order), C.byref(d), C.byref(f)) == 0
    return d.value, f.value


@pytest.mark.parametrize("scale", [1e-9, 1e-3, 1.0, 1e6, 1e13])
def test_sum_is_order_independent_and_accurate(scale):
    lib = load_emu()
    rng = np.random.default_rng(int(np.log10(scale)) + 20)
    # partials of very different magnitudes and both signs, like the per-CTA sums of a cancelling statistic
    v = rng.standard_normal(3000) * scale * 10.0 ** rng.uniform(-6, 0, 3000)
    ref = _det_sum(lib, v, np.arange(len(v)))
    for _ in range(5):
        assert _det_sum(lib, v, rng.permutation(len(v))) == ref          # bit for bit
    exact = float(sum(Fraction(float(x)) for x in v))
    # every partial is truncated to a multiple of 2^-56: the sum is low by less than n * 2^-56 (plus one rounding of
    # the result to double)
    assert -len(v) * 2.0 ** -56 - abs(exact) * 2.0 ** -52 <= ref[0] - exact <= abs(exact) * 2.0 ** -52
    # plain double accumulation of the same partials does depend on the order at these magnitudes
    if scale >= 1.0:
        sums = {float(np.sum(v[rng.permutation(len(v))])) for _ in range(20)} | {float(np.cumsum(v)[-1])}
        assert len(sums) > 1


def test_float_slots_negative_sums_and_saturation():
    lib = load_emu()
    v = np.array([-3.75, 1.5, -0.125, 2.0 ** -40, -2.0 ** -40])
    d, f = _det_sum(lib, v, np.arange(len(v)))
    assert d == -2.375 and f == np.float32(-2.375)
    # a non-finite or out-of-range partial saturates instead of poisoning the integer pair (documented behaviour)
    d, _ = _det_sum(lib, np.array([np.inf, 1.0]), np.arange(2))
    assert np.isfinite(d) and d > 1e16
    d, _ = _det_sum(lib, np.array([np.nan]), np.arange(1))
    assert np.isfinite(d)
