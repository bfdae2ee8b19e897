"""pt_detmath.h (the portable sin/cos/acos/pow the GPU path is specified to use) against glibc:
per-call deviation in ulps, and image-level effect of swapping it in (oracle mode 1 vs mode 0)."""
import numpy as np
import pytest

import oracle_api as O


def _ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


@pytest.mark.parametrize("fn,lo,hi,max_ulp", [
    (0, 0.0, 6.2832, 2), (1, 0.0, 6.2832, 2), (0, -100.0, 100.0, 2), (1, -100.0, 100.0, 2),
    (2, 0.0, 1.0, 2), (2, -1.0, 1.0, 2), (3, 0.0031308, 1.0, 1)])
def test_detmath_within_ulps_of_glibc(port, fn, lo, hi, max_ulp):
    rng = np.random.default_rng(fn)
    n = 1_000_000
    x = rng.uniform(lo, hi, n).astype(np.float32)
    y, z = np.zeros(n, np.float32), np.zeros(n, np.float32)
    e = np.float32(1.0) / np.float32(2.4)
    port.ptor_detmath(fn, x, y, n, e)
    port.ptor_detmath(fn + 10, x, z, n, e)
    assert _ulp_diff(y, z).max() <= max_ulp


def test_detmath_special_points(port):
    x = np.array([0.0, 1.0, -1.0, 0.5, -0.5], np.float32)
    y = np.zeros(5, np.float32)
    port.ptor_detmath(2, x, y, 5, 0.0)                      # acos
    assert y[1] == 0.0 and abs(float(y[0]) - np.pi / 2) < 2e-7 and abs(float(y[2]) - np.pi) < 4e-7
    x = np.array([0.0, 1.0], np.float32)
    y = np.zeros(2, np.float32)
    port.ptor_detmath(3, x, y, 2, np.float32(1 / 2.4))      # pow(0,y)=0, pow(1,y)=1
    assert y.tolist() == [0.0, 1.0]
    port.ptor_detmath(0, x, y, 2, 0.0)
    assert y[0] == 0.0
    port.ptor_detmath(1, x, y, 2, 0.0)
    assert y[0] == 1.0


@pytest.mark.parametrize("nee", [False, True])
def test_det_mode_close_to_libm_mode(port, cornell, nee):
    """Swapping glibc's sinf/cosf/acosf/powf for the portable versions perturbs directions by <= 2 ulp; a
    pixel can only change visibly if a discrete decision flips.  Tolerance (BASELINE north star): per channel
    |d| <= 1e-4 * max(|ref|, 1e-3); outliers are counted and bounded."""
    sc = O.OracleScene(port, cornell)
    p = O.make_params(port, 256, 256, 1, 16, nee=nee, importance=True)
    a = sc.render(p, math_mode=0)["accum"][..., :3]
    b = sc.render(p, math_mode=1)["accum"][..., :3]
    bad = (np.abs(a - b) > 1e-4 * np.maximum(np.abs(a), 1e-3)).any(axis=2)
    assert bad.mean() <= 2e-4, f"{bad.sum()} outlier pixels"
