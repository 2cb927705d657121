"""The DEVICE's pow (include/ppg_pow.h compiled by nvcc, `__fma_rn` where glibc's FMA build fuses) against the host libm —
bit for bit, through the C-ABI diagnostic `ppg_selftest_pow` (`-m gpu`)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _libm_pow(a, b):
    try:
        return math.pow(a, b)
    except OverflowError:  # libm returns +inf and sets ERANGE; CPython turns that into an exception
        return math.inf


def test_device_pow_is_bit_identical_to_libm():
    from predpreygrass_b200 import _lib

    L = _lib.load()
    rng = np.random.default_rng(11)
    n = 400_000
    u = lambda: rng.random(n)  # noqa: E731
    cases = {
        "speed ** 2": (0.5 + 1.5 * u(), np.full(n, 2.0)),
        "speed ** exponent": (0.5 + 1.5 * u(), 0.5 + 3.0 * u()),
        "(1 - p0) ** ratio": (u() * (1 - 1e-9) + 1e-9, 20.0 * u()),
        "wide": (np.ldexp(0.5 + u(), rng.integers(-1000, 1000, n).astype(np.int32)), (u() - 0.5) * 4.0),
        "near over/underflow": (0.5 + u(), (u() - 0.5) * 4000.0),
    }
    for name, (x, y) in cases.items():
        x, y = np.ascontiguousarray(x, np.float64), np.ascontiguousarray(y, np.float64)
        out = np.zeros(n)
        _lib.check(L.ppg_selftest_pow(x.ctypes.data, y.ctypes.data, out.ctypes.data, n, 0))
        ref = np.array([_libm_pow(a, b) for a, b in zip(x.tolist(), y.tolist())])  # CPython float pow = libm pow (numpy may use SIMD kernels)
        bad = np.nonzero(out.view(np.uint64) != ref.view(np.uint64))[0]
        assert bad.size == 0, (name, bad.size, x[bad[:3]], y[bad[:3]], out[bad[:3]], ref[bad[:3]])
