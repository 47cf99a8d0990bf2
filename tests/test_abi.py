"""The C-ABI boundary: both libraries load and export every symbol include/imc.h declares; the product
library fails loudly without a GPU (no CPU fallback); argument errors follow the documented convention."""
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
from mpimc_b200 import decks, driver, lib

HEADER = os.path.join(entry.ROOT, "include", "imc.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(imc_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_stage_entry_points():
    syms = declared_symbols()
    for s in ("imc_create", "imc_destroy", "imc_set_mesh", "imc_update", "imc_source", "imc_transport", "imc_clean",
              "imc_tally", "imc_energycheck", "imc_rw_table", "imc_step", "imc_reduce_buffer", "imc_get_particles",
              "imc_set_transport_tape", "imc_last_error"):
        assert s in syms


@pytest.mark.parametrize("which", ["cuda", "oracle"])
def test_library_exports_every_declared_symbol(built, which):
    path = entry.LIB if which == "cuda" else entry.ORACLE_LIB
    l = lib.ImcLib(path)
    for s in declared_symbols():
        assert hasattr(l.dll, s), f"{os.path.basename(path)} does not export {s}"
    assert {n for n, _, _ in lib.ABI} == set(declared_symbols()), "python binding and header disagree"
    assert l.backend == ("cuda-sm_100a" if which == "cuda" else "oracle-cpu")
    assert l.dll.imc_abi_version() == 1


def test_cuda_library_has_no_cpu_fallback(gpu_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.ImcError) as e:
        lib.Engine(lib.Config(precision=lib.F64, geometry=1, nx=10), gpu_lib)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_cuda_library_contains_sm100a_kernels(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", entry.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_argument_errors(oracle_lib):
    bad = [dict(precision=7), dict(geometry=3), dict(nx=0), dict(geometry=2, nx=4, ny=0), dict(bc=(5, 1, 1, 1)),
           dict(geometry=2, nx=2, ny=2, randomwalk=True)]
    for kw in bad:
        base = dict(precision=lib.F64, geometry=1, nx=10)
        base.update(kw)
        with pytest.raises(lib.ImcError) as e:
            lib.Engine(lib.Config(**base), oracle_lib)
        assert e.value.code == -1
    eng = lib.Engine(lib.Config(precision=lib.F64, geometry=1, nx=10), oracle_lib)
    with pytest.raises(lib.ImcError) as e:   # call out of order
        eng.update(0.1)
    assert e.value.code == -2
    with pytest.raises(ValueError):          # only REFLECT / VACUUM exist in the transport loop
        d = decks.suolson(); d["LEFTBC"] = "REFLECTIVE"
        driver.setup(d, oracle_lib)
