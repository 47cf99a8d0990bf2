"""Warp-cooperative device code checked on the CPU: csrc/imc_warp_reduce.cuh (the sequential EXACT-tally chains, among
them the experimental warp_seq_add_skip) is compiled for the host with __ballot_sync / __shfl_sync / __ffs emulated by 32
threads in lockstep (tests/warp_emu/warp_emu.h) and must reproduce a plain `v += record` loop bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(entry.ROOT, "build", "warp_emu", "libwarp_reduce_host.so")
SRC = os.path.join(HERE, "warp_emu", "warp_reduce_host.cpp")
DEPS = [SRC, os.path.join(HERE, "warp_emu", "warp_emu.h"), os.path.join(entry.CSRC, "imc_warp_reduce.cuh"), os.path.join(entry.CSRC, "imc_warp_runs.cuh"),
        os.path.join(entry.CSRC, "imc_num.h")]
T = {0: np.float16, 1: np.float32, 2: np.float64}


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-mf16c", "-pthread",
                        "-I" + entry.CSRC, "-I" + os.path.join(entry.ROOT, "include"), "-I" + os.path.join(HERE, "warp_emu"), "-o", OUT, SRC], check=True)
    dll = C.CDLL(OUT)
    up, dp = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    dll.warp_reduce_host.restype = C.c_double
    dll.warp_reduce_host.argtypes = [C.c_int, C.c_int, C.c_double, up, dp, C.c_longlong]
    dll.plain_chain.restype = C.c_double
    dll.plain_chain.argtypes = [C.c_int, C.c_double, up, dp, C.c_longlong]
    dll.plain_jl_sum.restype = C.c_double
    dll.plain_jl_sum.argtypes = [C.c_int, dp, C.c_longlong]

    def run(prec, v0, vals, keys=None):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        kp = None
        if keys is not None:
            keys = np.ascontiguousarray(keys, dtype=np.uint32); kp = keys.ctypes.data_as(up)
        a = [dll.warp_reduce_host(w, prec, v0, kp, vals.ctypes.data_as(dp), len(vals)) for w in (0, 1)]
        return a[0], a[1], dll.plain_chain(prec, v0, kp, vals.ctypes.data_as(dp), len(vals))

    def pairwise(prec, vals):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        return (dll.warp_reduce_host(2, prec, 0.0, None, vals.ctypes.data_as(dp), len(vals)), dll.plain_jl_sum(prec, vals.ctypes.data_as(dp), len(vals)))
    run.pairwise = pairwise
    dll.warp_runs_host.restype = C.c_longlong
    dll.warp_runs_host.argtypes = [C.c_int, C.POINTER(C.c_int), dp, C.c_int, dp]

    def runs(integer, cells, vals, ncell):
        cells = np.ascontiguousarray(cells, dtype=np.int32); vals = np.ascontiguousarray(vals, dtype=np.float64)
        sums = np.zeros(ncell)
        nd = dll.warp_runs_host(integer, cells.ctypes.data_as(C.POINTER(C.c_int)), vals.ctypes.data_as(dp), cells.shape[1], sums.ctypes.data_as(dp))
        return sums, nd
    run.runs = runs
    return run


def bits(x):
    return np.float64(x).tobytes()


@pytest.mark.parametrize("prec", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 64, 100, 257])
def test_chains_of_every_length(emu, prec, n):
    rng = np.random.default_rng(100 * prec + n)
    vals = (rng.normal(size=n) * 10.0 ** rng.integers(-4, 3, size=n)).astype(T[prec]).astype(np.float64)
    seq, skip, plain = emu(prec, 0.0, vals)
    assert bits(seq) == bits(plain) and bits(skip) == bits(plain)
    seq, skip, plain = emu(prec, float(T[prec](3.25)), vals)
    assert bits(seq) == bits(plain) and bits(skip) == bits(plain)


def test_stagnating_float16_sum(emu):
    """The case the skipping chain exists for: thousands of small deposits into a Float16 sum that soon stops moving."""
    rng = np.random.default_rng(7)
    vals = (rng.random(12000) ** 3 * 0.004).astype(np.float16).astype(np.float64)
    seq, skip, plain = emu(0, 0.0, vals)
    assert bits(seq) == bits(plain) and bits(skip) == bits(plain)
    assert 1.0 < plain < vals.sum() * 0.8                    # the Float16 sum has stagnated far below the exact one


@pytest.mark.parametrize("prec", [0, 1])
def test_special_values_and_wide_records(emu, prec):
    """Signed zeros, negative deposits, Inf / NaN, and MC_RW's Float64 records (bit 31 of the key) mixed in."""
    rng = np.random.default_rng(9 + prec)
    n = 300
    vals = (rng.normal(size=n) * 10.0 ** rng.integers(-6, 2, size=n)).astype(T[prec]).astype(np.float64)
    keys = np.where(rng.random(n) < 0.3, 0x80000000, 0).astype(np.uint32) | 5
    wide = (keys & 0x80000000) != 0
    vals[wide] = rng.normal(size=int(wide.sum())) * 1e-3     # Float64 values that are not representable in T
    for special in ([], [(10, -0.0)], [(0, -0.0), (1, -0.0)], [(50, np.inf)], [(50, np.inf), (200, -np.inf)], [(120, np.nan)]):
        v = vals.copy()
        for i, x in special:
            v[i] = x
        seq, skip, plain = emu(prec, 0.0, v, keys)
        assert bits(seq) == bits(plain) or (np.isnan(seq) and np.isnan(plain)), special
        assert bits(skip) == bits(plain) or (np.isnan(skip) and np.isnan(plain)), special
    seq, skip, plain = emu(prec, -0.0, np.array([-0.0, -0.0, 0.0, -0.0]))
    assert bits(seq) == bits(plain) == bits(skip)


@pytest.mark.parametrize("prec", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 2, 33, 1023, 1024, 1025, 2049, 4100])
def test_pairwise_sum_by_a_warp(emu, prec, n):
    """warp_jl_sum (PAIRWISE = TRUE: Julia's sum(vector), imc_transport.jl:202) against the plain recursion."""
    rng = np.random.default_rng(7 * prec + n)
    vals = (rng.random(n) * 10.0 ** rng.integers(-3, 2, size=n)).astype(T[prec]).astype(np.float64)
    got, want = emu.pairwise(prec, vals)
    assert bits(got) == bits(want)


@pytest.mark.parametrize("integer", [0, 1])
@pytest.mark.parametrize("per", [1, 4])
def test_runs_of_equal_cells_end_in_one_deposit(emu, integer, per):
    """imc_warp_runs.cuh (the census tally's ThreadRuns + warp_join_runs): 32 lanes x `per` consecutive particles with random
    runs of cells and dead particles in between — the per-cell sums equal the plain sums, and there is exactly one deposit per
    run of the list except for runs that start and end inside one lane's particles after its first run (deposited at once)
    and for runs broken by a lane without particles."""
    rng = np.random.default_rng(31 * per + integer)
    ncell = 6
    for trial in range(300):
        p_change = rng.choice([0.02, 0.1, 0.25, 0.6, 1.0])
        cells = np.empty((32, per), dtype=np.int32)
        cur = rng.integers(ncell)
        for i in range(32 * per):
            if rng.random() < p_change:
                cur = rng.integers(ncell)
            cells[i // per, i % per] = cur
        dead = rng.random((32, per)) < rng.choice([0.0, 0.1, 0.5])
        if trial % 7 == 0:
            dead[rng.integers(32)] = True                      # a lane without particles
        if trial % 11 == 0:
            dead[:] = True                                     # nothing alive at all
        vals = rng.integers(1, 1000, size=(32, per)).astype(np.float64)
        sums, nd = emu.runs(integer, np.where(dead, -1, cells), vals, ncell)
        want = np.zeros(ncell)
        np.add.at(want, cells[~dead], vals[~dead])
        assert np.array_equal(sums, want), (trial, per, integer)
        alive_cells = cells[~dead]
        runs_in_list = 0 if alive_cells.size == 0 else 1 + int(np.count_nonzero(alive_cells[1:] != alive_cells[:-1]))
        assert nd >= runs_in_list                              # never fewer deposits than runs ...
        if not dead.any():
            assert nd <= runs_in_list + 1                      # ... and no more (lane 0's first run cannot join a previous warp)
