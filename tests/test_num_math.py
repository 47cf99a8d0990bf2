"""Number type, deterministic math, Philox and Julia-sum helpers (shared by the kernels and the oracle),
checked against independent implementations: numpy float16, numpy/libm Float64, Random123's published
known-answer vectors, and a pure-Python restatement of Julia's pairwise sum."""
import ctypes as C

import numpy as np
import pytest

from mpimc_b200 import lib

DP = C.POINTER(C.c_double)


def math_eval(olib, fn, prec, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    yp = None if y is None else np.ascontiguousarray(y, dtype=np.float64).ctypes.data_as(DP)
    f = olib.dll.imc_oracle_math_eval
    f.restype = C.c_int
    f.argtypes = [C.c_int32, C.c_int32, DP, DP, DP, C.c_int64]
    assert f(fn, prec, x.ctypes.data_as(DP), yp, out.ctypes.data_as(DP), x.size) == 0
    return out


def test_float16_rounding_matches_numpy(oracle_lib):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=20000) * 10.0 ** rng.integers(-9, 5, size=20000), [0.0, 65504.0, 65519.9, 65520.0, 1e-8, 5.96e-8, 2049.0, 50000.0]])
    got = math_eval(oracle_lib, 7, lib.F16, x)
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).astype(np.float64)
    assert np.array_equal(got, want)
    assert got[-1] == 49984.0 and got[-2] == 2048.0
    # Float32 rounding
    assert np.array_equal(math_eval(oracle_lib, 7, lib.F32, x), x.astype(np.float32).astype(np.float64))


@pytest.mark.parametrize("op,fn", [("mul", 9), ("add", 10), ("div", 11)])
def test_float16_arithmetic_rounds_after_every_operation(oracle_lib, op, fn):
    rng = np.random.default_rng(1)
    a = (rng.normal(size=50000) * 10.0 ** rng.integers(-3, 3, size=50000)).astype(np.float16)
    b = (rng.normal(size=50000) * 10.0 ** rng.integers(-3, 3, size=50000)).astype(np.float16)
    b[b == 0] = np.float16(1.0)
    with np.errstate(all="ignore"):
        want = {"mul": a * b, "add": a + b, "div": a / b}[op].astype(np.float64)
    got = math_eval(oracle_lib, fn, lib.F16, a.astype(np.float64), b.astype(np.float64))
    assert np.array_equal(got, want, equal_nan=True)


def ulp_err(got, want, dtype):
    want_t = want.astype(dtype)
    ulp = np.abs(np.spacing(want_t)).astype(np.float64)
    return np.max(np.abs(got - want) / ulp)


@pytest.mark.parametrize("name,fn,ref,lo,hi,tol", [
    ("exp", 0, np.exp, -40.0, 10.0, 1.0), ("exp_small", 0, np.exp, -1.0, 0.0, 1.0),
    ("expm1", 1, np.expm1, -30.0, 2.0, 2.0), ("expm1_small", 1, np.expm1, -1e-2, 0.0, 1.0),
    ("log", 2, np.log, 1e-30, 1.0, 1.0), ("log_wide", 2, np.log, 1e-3, 1e6, 1.0),
    ("sin", 3, np.sin, -7.0, 7.0, 1.5), ("cos", 4, np.cos, -7.0, 7.0, 1.5), ("sqrt", 8, np.sqrt, 0.0, 100.0, 0.5),
])
def test_deterministic_math_accuracy(oracle_lib, name, fn, ref, lo, hi, tol):
    """imc_math.h against numpy (glibc) — < 1 ulp on the transport's ranges, <= 2 ulp for expm1 at large |x|."""
    rng = np.random.default_rng(2)
    x64 = rng.uniform(lo, hi, size=200000)
    assert ulp_err(math_eval(oracle_lib, fn, lib.F64, x64), ref(x64), np.float64) <= tol + 0.01
    x32 = x64.astype(np.float32).astype(np.float64)
    assert ulp_err(math_eval(oracle_lib, fn, lib.F32, x32), ref(x32), np.float32) <= tol + 0.01
    # Float16: Julia computes in Float32 and rounds once
    x16 = x64.astype(np.float16).astype(np.float64)
    x16 = x16[np.isfinite(x16)]
    with np.errstate(all="ignore"):
        want16 = ref(x16.astype(np.float32)).astype(np.float16).astype(np.float64)
    got16 = math_eval(oracle_lib, fn, lib.F16, x16)
    ok = np.isfinite(want16)
    assert np.mean(got16[ok] == want16[ok]) > 0.999   # identical except rare double-rounding ties


def test_atan2_and_pow(oracle_lib):
    rng = np.random.default_rng(3)
    y, x = rng.normal(size=100000), rng.normal(size=100000)
    assert ulp_err(math_eval(oracle_lib, 5, lib.F64, y, x), np.arctan2(y, x), np.float64) <= 2.0
    assert math_eval(oracle_lib, 5, lib.F64, [0.0, -0.0, 1.0], [-1.0, -1.0, 0.0]).tolist() == [np.pi, -np.pi, np.pi / 2]
    b = rng.uniform(0.01, 3.0, size=1000)
    for p, exact in ((0.0, True), (1.0, True), (0.25, False), (-3.0, False)):
        got = math_eval(oracle_lib, 6, lib.F64, b, np.full_like(b, p))
        if exact:
            assert np.array_equal(got, b ** p)
        else:
            assert ulp_err(got, b ** p, np.float64) <= 2.0


def test_philox_known_answers(oracle_lib):
    """Random123 kat_vectors for philox4x32-10."""
    f = oracle_lib.dll.imc_oracle_philox
    f.restype = None
    U4, U2 = C.c_uint32 * 4, C.c_uint32 * 2
    f.argtypes = [U4, U2, U4]
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        out = U4()
        f(U4(*ctr), U2(*key), out)
        assert tuple(out) == want


def test_draws_follow_julia_conventions(oracle_lib):
    f = oracle_lib.dll.imc_oracle_draws
    f.restype = None
    f.argtypes = [C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, DP, C.c_int64]
    n = 20000
    for prec, bits in ((lib.F16, 11), (lib.F32, 24), (lib.F64, 53)):
        u = np.empty(n); e = np.empty(n)
        f(prec, 1234, 77, 3, 1, 0, u.ctypes.data_as(DP), n)
        f(prec, 1234, 77, 3, 1, 1, e.ctypes.data_as(DP), n)
        assert u.min() >= 0 and u.max() < 1
        assert np.array_equal(u * 2.0 ** bits, np.round(u * 2.0 ** bits))   # rand(T) is a multiple of 2^-bits
        assert abs(u.mean() - 0.5) < 0.01 and e.min() >= 0 and abs(e.mean() - 1.0) < 0.03
        u2 = np.empty(n)
        f(prec, 1234, 78, 3, 1, 0, u2.ctypes.data_as(DP), n)   # another particle id: another stream
        assert not np.array_equal(u, u2)
        u3 = np.empty(n)
        f(prec, 1234, 77, 3, 1, 0, u3.ctypes.data_as(DP), n)   # counter-based: reproducible
        assert np.array_equal(u, u3)


def julia_sum(a):
    """Base.sum over a Vector (mapreduce_impl, block size 1024) restated in Python."""
    def rec(first, last):
        if first == last:
            return a[first]
        if last - first < 1024:
            v = a[first] + a[first + 1]
            for i in range(first + 2, last + 1):
                v = v + a[i]
            return v
        mid = first + ((last - first) >> 1)
        return rec(first, mid) + rec(mid + 1, last)
    return rec(0, len(a) - 1) if len(a) else a.dtype.type(0)


@pytest.mark.parametrize("n", [1, 2, 15, 16, 1023, 1024, 1025, 2048, 2049, 2050, 4099, 10000])
def test_julia_pairwise_sum(oracle_lib, n):
    rng = np.random.default_rng(n)
    f = oracle_lib.dll.imc_oracle_jl_sum
    f.restype = C.c_double
    f.argtypes = [C.c_int32, DP, C.c_int64]
    for prec, T in ((lib.F32, np.float32), (lib.F64, np.float64)):
        a = (rng.random(n) * 10.0 ** rng.integers(-4, 4, size=n)).astype(T)
        got = f(prec, a.astype(np.float64).ctypes.data_as(DP), n)
        assert got == float(julia_sum(a))


def _eval(oracle_lib, fn, prec, x, y=None):
    f = oracle_lib.dll.imc_oracle_math_eval
    f.restype = C.c_int
    f.argtypes = [C.c_int32, C.c_int32, DP, DP, DP, C.c_int64]
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    yp = None if y is None else np.ascontiguousarray(y, dtype=np.float64).ctypes.data_as(DP)
    assert f(fn, prec, x.ctypes.data_as(DP), yp, out.ctypes.data_as(DP), x.size) == 0
    return out


def test_exponential_sampler_transform(oracle_lib):
    """dm::neglog_unit_f, the division-free -log(u) behind randexp(Float16 / Float32): within 2 ulp of -log(u) for every
    kind of uniform the generator can produce, exactly 0 at u = 1, finite at the smallest u, never negative."""
    rng = np.random.default_rng(5)
    w = np.concatenate([rng.integers(0, 2 ** 32, 400000, dtype=np.uint64), [0, 1, 2, 2 ** 31, 2 ** 32 - 2, 2 ** 32 - 1],
                        2 ** 32 - 1 - rng.integers(0, 4096, 2000, dtype=np.uint64)]).astype(np.float64)
    u = np.minimum((w.astype(np.float32) * np.float32(2.3283064365386963e-10) + np.float32(1.1641532182693481e-10)).astype(np.float32), np.float32(1.0))
    got = _eval(oracle_lib, 12, lib.F32, u.astype(np.float64))
    want = -np.log(u.astype(np.float64))
    assert np.all(got >= 0) and np.all(np.isfinite(got))
    assert got[u == 1.0].max() == 0.0
    nz = want > 0
    ulp = np.spacing(want[nz].astype(np.float32)).astype(np.float64)
    assert np.max(np.abs(got[nz] - want[nz]) / ulp) <= 2.0
    assert 22.0 < got.max() < 23.0       # u = 2^-33


@pytest.mark.parametrize("prec,fn", [(lib.F64, 13), (lib.F32, 14)])
def test_fused_exp_expm1_equals_the_separate_functions(oracle_lib, prec, fn):
    """exp_expm1 (one range reduction, and none at all when |x| < ln2/2 — the tracking loop's usual argument) returns
    bit for bit what exp and expm1 return separately."""
    rng = np.random.default_rng(6)
    x = np.concatenate([-rng.uniform(0, 0.5, 200000), -rng.uniform(0, 1e-3, 100000), rng.uniform(-40, 40, 100000), -10.0 ** rng.uniform(-45, 1, 50000),
                        [0.0, -0.0, 0.34657, -0.34657, 0.3466, -0.3466, 1e-300, -1e-300, 88.0, -17.0, -37.0, 709.0, -745.0]])
    if prec == lib.F32:
        x = x.astype(np.float32).astype(np.float64)
    e = _eval(oracle_lib, fn, prec, x, np.zeros_like(x))
    m = _eval(oracle_lib, fn, prec, x, np.ones_like(x))
    assert np.array_equal(e, _eval(oracle_lib, 0, prec, x), equal_nan=True)
    assert np.array_equal(m, _eval(oracle_lib, 1, prec, x), equal_nan=True)
