"""imc_checkpoint (include/imc.h): the restart point inside the library.  A run that is saved, advanced, restored and
advanced again must repeat itself bit for bit — particles, fields, every scalar of the stage statistics — on the oracle
(CPU) and on the CUDA engine; bench.py relies on it to time the resident and the host-buffer path over the same steps."""
import copy

import numpy as np
import pytest

from mpimc_b200 import decks, driver, lib

DECKS = {
    "suolson-f32": lambda: decks.suolson(precision="FLOAT32", n_input=1500, n_max=12000),
    "crookedpipe-f64": lambda: decks.crooked_pipe(precision="FLOAT64", n_input=2000, n_max=15000),
    "marshak-rw-f32": lambda: decks.marshak(precision="FLOAT32", n_cells=64, nonuniform=True, randomwalk="TRUE", n_input=2000, n_max=15000),
}


def _state(sim):
    p, ids = sim.engine.particles()
    return {"p": p.copy(), "ids": ids.copy(), **{f: sim.engine.field(f).copy() for f in ("temp", "fleck", "matenergydens", "radenergydens", "energydep")}}


def _scalars(rec):
    return {k: rec[k] for k in ("source", "transport", "tally", "energy")}


def _strip(d):   # timing fields differ run to run
    d = copy.deepcopy(d)
    d["transport"].pop("kernel_ms", None); d["transport"].pop("variant", None)
    return d


def _roundtrip(library, name, **kw):
    sim = driver.setup(DECKS[name](), library, **kw)
    for _ in range(2):
        sim.advance()
    sim.engine.checkpoint("save")
    sv = copy.deepcopy(sim.simvars)
    first = [_strip(_scalars(sim.advance())) for _ in range(3)]
    s1 = _state(sim)
    sim.engine.checkpoint("restore")
    sim.simvars = copy.deepcopy(sv)
    second = [_strip(_scalars(sim.advance())) for _ in range(3)]
    s2 = _state(sim)
    assert first == second, name
    for k in s1:
        assert np.array_equal(s1[k], s2[k], equal_nan=True), (name, k)
    # a second restore works too, and drop makes the next restore an error
    sim.engine.checkpoint("restore")
    sim.engine.checkpoint("drop")
    with pytest.raises(lib.ImcError):
        sim.engine.checkpoint("restore")


@pytest.mark.parametrize("name", sorted(DECKS))
def test_oracle_checkpoint_roundtrip(oracle_lib, name):
    _roundtrip(oracle_lib, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(DECKS))
@pytest.mark.parametrize("mode", ["exact", "fixed"])
def test_cuda_checkpoint_roundtrip(gpu_lib, name, mode):
    # EXACT and FIXED tallies are deterministic, so the repeated steps must agree in every bit
    _roundtrip(gpu_lib, name, tally_mode={"exact": lib.TALLY_EXACT, "fixed": lib.TALLY_FIXED}[mode])
