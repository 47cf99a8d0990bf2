"""csrc/imc_sortnet.h (the register sorting network behind Utilities.sorter on the device) on the CPU: every size the
kernels use sorts all 2^N zero-one inputs (0-1 principle => it sorts everything), and random doubles with duplicates,
zeros and infinities come out like numpy's sort."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(entry.ROOT, "build", "sortnet", "libsortnet_host.so")
SRC = os.path.join(HERE, "sortnet", "sortnet_host.cpp")
DEPS = [SRC, os.path.join(entry.CSRC, "imc_sortnet.h")]


@pytest.fixture(scope="module")
def dll():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I" + entry.CSRC, "-o", OUT, SRC], check=True)
    d = C.CDLL(OUT)
    d.sortnet_ok.restype = C.c_int; d.sortnet_ok.argtypes = [C.c_int]
    d.sortnet_sort13.argtypes = [C.POINTER(C.c_double)]
    return d


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 16])
def test_network_sorts_every_zero_one_input(dll, n):
    assert dll.sortnet_ok(n) == 1


def test_network_matches_numpy_sort(dll):
    rng = np.random.default_rng(5)
    for _ in range(2000):
        v = rng.choice([0.0, 1.0, 1e-300, 2.5, -3.0, np.inf, 7.0, 1e300, 0.1], size=13) * rng.choice([1.0, 1.0, 0.5], size=13)
        w = v.copy()
        dll.sortnet_sort13(w.ctypes.data_as(C.POINTER(C.c_double)))
        assert np.array_equal(w, np.sort(v))
