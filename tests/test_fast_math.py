"""The branch-free FP64 elementary functions of csrc/crcl_common.cuh (namespace fm), compiled for the CPU with a 20-bit
model of the MUFU seeds, against numpy / mpmath: errors in units of the last place over the argument ranges the
surfaces produce, and the special values the kernels rely on (NaN propagation, acos at +-1, exp far out of range)."""
import ctypes

import numpy as np
import pytest


def _fm(H, func, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(x if y is None else y, dtype=np.float64)
    out = np.empty_like(x)
    P = ctypes.POINTER(ctypes.c_double)
    H.hh_fm(ctypes.c_int(func), x.ctypes.data_as(P), y.ctypes.data_as(P), ctypes.c_int(x.size), out.ctypes.data_as(P))
    return out


def _ulps(got, ref_ld):
    """|got - ref| in ulps of ref; ref in extended precision"""
    ref = ref_ld.astype(np.float64)
    return np.abs((got.astype(np.longdouble) - ref_ld) / np.spacing(np.abs(ref)).astype(np.longdouble)).max()


RNG = np.random.default_rng(7)
N = 400_000


def test_rcp_div_rsqrt_sqrt(host_harness):
    x = np.exp(RNG.uniform(-60, 60, N)) * RNG.choice([-1.0, 1.0], N)
    xl = x.astype(np.longdouble)
    assert _ulps(_fm(host_harness, 0, x), 1 / xl) <= 1.0
    a = RNG.normal(0, 10, N)
    assert _ulps(_fm(host_harness, 1, a, x), a.astype(np.longdouble) / xl) <= 1.0
    xp = np.abs(x)
    xpl = xp.astype(np.longdouble)
    assert _ulps(_fm(host_harness, 2, xp), 1 / np.sqrt(xpl)) <= 1.5
    assert _ulps(_fm(host_harness, 3, xp), np.sqrt(xpl)) <= 1.0
    assert _ulps(_fm(host_harness, 6, xp), np.sqrt(xpl)) <= 1.0


def test_exp(host_harness):
    x = np.concatenate([RNG.uniform(-707, 709, N), RNG.uniform(-40, 5, N), RNG.normal(0, 1e-3, 1000), [0.0]])
    got = _fm(host_harness, 4, x)
    assert _ulps(got, np.exp(x.astype(np.longdouble))) <= 1.5
    assert got[-1] == 1.0
    # far out of range: 0 and inf as the library gives them; NaN propagates
    far = _fm(host_harness, 4, np.array([-800.0, -1e6, 800.0, np.nan]))
    assert far[0] == 0.0 and far[1] == 0.0 and far[2] == np.inf and np.isnan(far[3])
    # beyond 2^31 ln2 the integer part no longer fits the low word of the magic-number sum (BKMP2's singlet curve
    # evaluates exp(-2e12) at R = 30 a0): still ~0 / huge, for every magnitude up to inf
    big = np.concatenate([10.0 ** np.arange(3.1, 300, 1.7), [np.inf, 2.0e12, 1.49e9, 1.5e9, 2 ** 31 * 0.7, 2 ** 32 * 0.7]])
    lo, hi = _fm(host_harness, 4, -big), _fm(host_harness, 4, big)
    assert (lo == 0.0).all() and (hi == np.inf).all()
    # the denormal range is flushed to zero, the last normal results are exact to the usual bound
    edge = _fm(host_harness, 4, np.array([-707.3, -708.5, -720.0, -745.0]))
    assert abs(edge[0] / np.exp(-707.3) - 1) < 1e-15 and (edge[1:] == 0.0).all()


def test_acos(host_harness):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    x = np.concatenate([RNG.uniform(-1, 1, N), 1 - np.exp(RNG.uniform(-36, 0, 20000)), -1 + np.exp(RNG.uniform(-36, 0, 20000)),
                        [0.5, -0.5, 0.0, np.nextafter(0.5, 1), np.nextafter(-0.5, -1)]])
    got = _fm(host_harness, 5, x)
    # numpy's longdouble arccos is accurate to ~1e-19 here
    assert _ulps(got, np.arccos(x.astype(np.longdouble))) <= 2.0
    # spot values in 40-digit arithmetic
    for v in (-1 / 3, 0.123456789, -0.87, 0.999999, -0.999999):
        g = _fm(host_harness, 5, np.array([v]))[0]
        assert abs(g - float(mp.acos(mp.mpf(v)))) <= 2 * np.spacing(g)
    end = _fm(host_harness, 5, np.array([1.0, -1.0, np.nan]))
    assert end[0] == 0.0 and end[1] == np.pi and np.isnan(end[2])


def test_log_pow(host_harness):
    x = np.concatenate([np.exp(RNG.uniform(-300, 300, N)), RNG.uniform(0.5, 2.0, N), 1 + RNG.normal(0, 1e-6, 2000), [1.0]])
    got = _fm(host_harness, 7, x)
    ref = np.log(x.astype(np.longdouble))
    nz = ref != 0
    assert _ulps(got[nz], ref[nz]) <= 2.0
    assert got[-1] == 0.0
    # x^y with |y log x| < 10: a few ulp, growing with |y log x|
    xb = RNG.uniform(0.3, 30.0, N)
    yb = RNG.uniform(-2.5, 2.5, N)
    gp = _fm(host_harness, 8, xb, yb)
    rp = np.exp(yb.astype(np.longdouble) * np.log(xb.astype(np.longdouble)))
    assert _ulps(gp, rp) <= 20.0
    assert np.abs(gp / rp.astype(np.float64) - 1).max() < 3e-15
    z = _fm(host_harness, 9, np.array([0.0, 4.0, 1e-200]))
    assert z[0] == 0.0 and z[1] == 2.0 and abs(z[2] - 1e-100) < 1e-115
