"""Device-only arithmetic of the pair kernels against libm (EmDeeX_math_probe): the reciprocals that replace the reference's
divisions (rcp.approx.f64 + cubic refinement), the typed kernel's exp of a non-positive argument (constant-bank
coefficients, no special-case branch) and its Abramowitz-Stegun erfc (reference src/math.f90:685-691). The GPU test is
the one that exercises the device code; the emulator twin (same source compiled for the host) pins the algorithm."""
import math

import numpy as np
import pytest

import common as cm


def _probe(lib, what, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib.EmDeeX_math_probe(what, x.size, x.ctypes.data, out.ctypes.data)
    return out


def _uerfc_ref(x):
    a1, a2, a3, a4, a5, p = 0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429, 0.327591100
    t = 1.0 / (1.0 + p * x)
    return t * (a1 + t * (a2 + t * (a3 + t * (a4 + t * a5)))) * np.exp(-x * x)


def check_device_math(lib):
    rng = np.random.default_rng(7)
    # reciprocals: the whole range a squared distance or a model argument can take, plus the typical r^2 band
    a = np.concatenate([10.0 ** rng.uniform(-290, 290, 20000), rng.uniform(0.5, 10.0, 20000), 2.0 ** np.arange(-1000, 1000, 7.0)])
    for what in (0, 1):
        y = _probe(lib, what, a)
        assert np.all(np.isfinite(y))
        assert np.max(np.abs(y * a - 1.0)) < 4.5e-16, what   # within 2 ulp of 1/a
    # coincident atoms: documented deviation -- no trap, no hang; the value is NaN or Inf (deviation ii)
    z = _probe(lib, 0, np.array([0.0]))
    assert not np.isfinite(z[0])
    # exp of a non-positive argument
    t = np.concatenate([-rng.uniform(0.0, 40.0, 40000), -rng.uniform(0.0, 690.0, 20000), [0.0, -1e-300, -1e-17, -690.0]])
    y = _probe(lib, 2, t)
    ref = np.exp(t)
    assert np.max(np.abs(y - ref) / ref) < 4.5e-16
    far = _probe(lib, 2, np.array([-700.0, -800.0, -1.0e4, -1.0e8, -1.0e300]))
    assert np.all(np.isfinite(far)) and np.all(far >= 0.0) and np.all(far < 1e-290)   # clamped exponent: tiny, never garbage
    # the reference's erfc: typed-kernel form and generic-kernel form against the formula in numpy
    x = np.concatenate([rng.uniform(0.0, 6.0, 40000), [0.0, 1e-12, 26.0]])
    ref = _uerfc_ref(x)
    for what in (3, 4):
        y = _probe(lib, what, x)
        tol = 2e-15 * np.maximum(ref, 1e-300)
        assert np.all(np.abs(y - ref) <= tol + 1e-300), what


def test_device_math_on_emulator():
    check_device_math(cm.emulated())


@pytest.mark.gpu
def test_device_math_on_gpu():
    check_device_math(cm.product())
