"""include/eucl_detmath.h (the libm shared by the CUDA path and the `det` oracle) against
long-double references and glibc: every function must stay a faithful libm (< 1 ulp; atan2 < 2)."""
import ctypes as C

import numpy as np
import pytest

dp = C.POINTER(C.c_double)
RNG = np.random.default_rng(20261018)
N = 200_000


def call_unary(oracle, fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    oracle.lib("det").oracle_detmath_unary(fn, x.ctypes.data_as(dp), out.ctypes.data_as(dp), len(x))
    return out


def max_ulp(got, ref_ld):
    refd = ref_ld.astype(np.float64)
    ok = np.isfinite(refd)
    with np.errstate(invalid="ignore"):
        err = np.abs(got[ok].astype(np.longdouble) - ref_ld[ok]) / np.spacing(np.maximum(np.abs(refd[ok]), 1e-300))
    return float(err.max())


def unit_inputs():
    u = RNG.uniform(-1, 1, N)
    small = RNG.uniform(-1, 1, N) * 10.0 ** RNG.uniform(-12, 0, N)
    edge = np.concatenate([1 - 10.0 ** RNG.uniform(-16, -1, N // 4), -1 + 10.0 ** RNG.uniform(-16, -1, N // 4),
                           [1.0, -1.0, 0.0, -0.0, 0.5, -0.5, 0.975, -0.975]])
    return np.concatenate([u, small, edge])


@pytest.mark.parametrize("fn,ref", [(0, np.arccos), (1, np.arcsin)])
def test_acos_asin(oracle, fn, ref):
    x = unit_inputs()
    got = call_unary(oracle, fn, x)
    assert max_ulp(got, ref(x.astype(np.longdouble))) < 1.0
    bad = call_unary(oracle, fn, np.array([1.0000000000000002, -1.0000000000000002, 2.0, np.nan]))
    assert np.isnan(bad).all()  # the renderer relies on NaN for |x| > 1 (angle_between -> 0, Fresnel TIR -> 1)


@pytest.mark.parametrize("fn,ref", [(2, np.sin), (3, np.cos)])
def test_sin_cos(oracle, fn, ref):
    x = np.concatenate([RNG.uniform(-np.pi, np.pi, N), RNG.uniform(-100, 100, N), RNG.uniform(-1e5, 1e5, N // 4),
                        RNG.uniform(-1, 1, N) * 10.0 ** RNG.uniform(-12, 0, N),
                        np.pi / 2 * np.arange(-40, 41) + RNG.uniform(-1e-9, 1e-9, 81), [0.0, -0.0, np.pi / 4, np.pi / 2, np.pi]])
    assert max_ulp(call_unary(oracle, fn, x), ref(x.astype(np.longdouble))) < 1.0
    assert np.isnan(call_unary(oracle, fn, np.array([np.inf, -np.inf, np.nan]))).all()


def test_atan_atan2(oracle):
    x = np.concatenate([RNG.uniform(-5, 5, N), RNG.uniform(-1, 1, N) * 10.0 ** RNG.uniform(-10, 10, N),
                        [0.0, -0.0, 1.0, -1.0, 0.4375, 0.6875, 1.1875, 2.4375]])
    assert max_ulp(call_unary(oracle, 4, x), np.arctan(x.astype(np.longdouble))) < 1.0
    y = np.concatenate([RNG.uniform(-1, 1, N), RNG.uniform(-1, 1, N) * 10.0 ** RNG.uniform(-10, 10, N)])
    xx = np.concatenate([RNG.uniform(-1, 1, N), RNG.uniform(-1, 1, N) * 10.0 ** RNG.uniform(-10, 10, N)])
    out = np.empty_like(y)
    oracle.lib("det").oracle_detmath_atan2(y.ctypes.data_as(dp), xx.ctypes.data_as(dp), out.ctypes.data_as(dp), len(y))
    assert max_ulp(out, np.arctan2(y.astype(np.longdouble), xx.astype(np.longdouble))) < 2.0
    # special cases agree with glibc bit for bit (signs of zero included)
    ys = np.array([0.0, -0.0, 0.0, -0.0, 1, 1, -1, -1, np.inf, -np.inf, np.inf, 1.0, 0.0, np.nan])
    xs = np.array([1.0, 1.0, -1.0, -1.0, 0.0, -0.0, 0.0, -0.0, np.inf, np.inf, -np.inf, -np.inf, 0.0, 1.0])
    o = np.empty_like(ys)
    oracle.lib("det").oracle_detmath_atan2(ys.ctypes.data_as(dp), xs.ctypes.data_as(dp), o.ctypes.data_as(dp), len(ys))
    want = np.arctan2(ys, xs)
    assert np.array_equal(o, want, equal_nan=True) and np.array_equal(np.signbit(o), np.signbit(want))


def test_both_oracle_builds_report_their_libm(oracle):
    assert oracle.lib("det").oracle_uses_detmath() == 1
    assert oracle.lib("glibc").oracle_uses_detmath() == 0
