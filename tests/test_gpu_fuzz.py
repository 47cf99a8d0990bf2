"""Randomised parity: small random 2-D and 1-D decks (mesh shape and grading, boundary conditions, opacities and their
temperature powers, scattering, sources, energy scales, distance scale, speed of light, time step, precision, PAIRWISE)
run three steps on the CUDA engine (EXACT tallies, warp-refill or static schedule) and on the oracle: particles, outcome
counts and every field must be bit-identical.  Seeds are fixed, so a failure names a reproducible deck."""
import numpy as np
import pytest

from mpimc_b200 import decks, driver, lib
from test_gpu_parity import FIELDS_EXACT, FIELDS_TALLIED, assert_step_parity, run_pair

pytestmark = pytest.mark.gpu


def random_2d(rng, precision):
    f16 = precision == "FLOAT16"
    nx, ny = int(rng.integers(1, 9)), int(rng.integers(1, 9))
    xs = np.concatenate([[0.0], np.cumsum(rng.uniform(0.02, 0.4, nx))]).round(6)
    ys = np.concatenate([[0.0], np.cumsum(rng.uniform(0.02, 0.4, ny))]).round(6)
    bcs = tuple(rng.choice(["REFLECT", "VACUUM"], 4))
    scales = (1024.0,) if f16 else tuple(sorted(set(rng.choice([4.0, 1.0, 0.5, 64.0], int(rng.integers(1, 4))).tolist()), reverse=True))
    d = decks.small_2d(precision=precision, n_input=int(rng.integers(200, 1500)), n_max=60000, bcs=bcs, energyscales=scales,
                       pairwise=str(rng.choice(["TRUE", "FALSE"])), seed=int(rng.integers(1, 10**6)))
    d["XMESHNODES"], d["YMESHNODES"] = xs, ys
    d["T_SURFACE_REGS"] = [(float(xs[-1]), float(xs[-1]), float(ys[-1]), float(ys[-1]))]
    d["T_SURFACE_VALS"] = [tuple(float(v) for v in rng.choice([0.0, 0.5, 1.0], 4))]
    d["SIGMA_A_VALS"] = [repr(float(rng.choice([0.2, 1.0, 5.0, 40.0])))]
    d["SIGMA_A_POWERS"] = [repr(float(rng.choice([0.0, -3.0, 1.0])))] if not f16 else ["0.0"]
    d["SIGMA_S_VALS"] = [repr(float(rng.choice([0.0, 0.5, 10.0])))]
    d["RADSOURCE_VALS"] = [repr(float(rng.choice([0.0, 1.0])))]
    d["T_INIT"] = repr(float(rng.choice([0.1, 0.5, 1.0])))
    d["DT"] = repr(float(rng.choice([0.001, 0.01, 0.05])))
    d["PHYS_C"] = repr(float(rng.choice([1.0, 299.70])))
    d["DISTANCESCALE"] = repr(float(rng.choice([1.0, 1.0, 2.0, 3.0])))
    d["ALPHA"] = repr(float(rng.choice([0.5, 1.0])))
    return d


def random_1d(rng, precision):
    f16 = precision == "FLOAT16"
    kind = rng.choice(["suolson", "nonuniform"])
    if kind == "suolson":
        d = decks.suolson(precision=precision, n_input=int(rng.integers(300, 3000)), n_max=30000, pairwise=str(rng.choice(["TRUE", "FALSE"])))
    else:
        scales = (1024.0,) if f16 else (64.0, 4.0, 1.0, 0.5)
        d = decks.nonuniform_1d(precision=precision, n_input=int(rng.integers(300, 3000)), energyscales=scales, pairwise=str(rng.choice(["TRUE", "FALSE"])))
        d["LEFTBC"], d["RIGHTBC"] = str(rng.choice(["REFLECT", "VACUUM"])), str(rng.choice(["REFLECT", "VACUUM"]))
        d["SIGMA_A_VALS"] = [repr(float(rng.choice([0.2, 2.0, 20.0])))]
        d["SIGMA_S_VALS"] = [repr(float(rng.choice([0.0, 1.0])))]
    d["SEED"] = str(int(rng.integers(1, 10**6)))
    d["DISTANCESCALE"] = repr(float(rng.choice([1.0, 1.0, 2.0])))
    return d


@pytest.mark.parametrize("seed", range(96))
def test_random_decks_bit_exact(gpu_lib, oracle_lib, seed):
    rng = np.random.default_rng(1000 + seed)
    precision = ["FLOAT64", "FLOAT32", "FLOAT16"][seed % 3]
    inputs = random_2d(rng, precision) if seed % 4 != 3 else random_1d(rng, precision)
    track = [lib.TRACK_REFILL, lib.TRACK_HISTORY][seed % 2]
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=3, sync=False, tally_mode=lib.TALLY_EXACT, track_mode=track)
    assert_step_parity(a, b, out, precision)
    for ra, rb in out:
        assert ra["transport"]["lostenergy"] == rb["transport"]["lostenergy"]
        assert ra["tally"] == rb["tally"] or (np.isnan(ra["tally"]["max_temp"]) and np.isnan(rb["tally"]["max_temp"]))
    for name in FIELDS_EXACT + FIELDS_TALLIED + ("bee",):
        fa, fb = a.engine.field(name), b.engine.field(name)
        assert np.array_equal(fa, fb, equal_nan=True), name


@pytest.mark.parametrize("precision,track", [("FLOAT32", lib.TRACK_REFILL), ("FLOAT64", lib.TRACK_REFILL), ("FLOAT32", lib.TRACK_HISTORY)])
def test_mid_scale_crooked_pipe_bit_exact(gpu_lib, oracle_lib, precision, track):
    """1.2 million particles on a 160 x 160 crooked pipe (every SM busy, the refill queue contended, global-memory
    tallies): still bit-identical to the oracle in every particle and, with EXACT tallies, every field."""
    inputs = decks.crooked_pipe(precision=precision, n_input=1_200_000, n_max=4_000_000, cellmin=1, mesh_cells=(160, 160), pairwise="FALSE")
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=2, sync=False, tally_mode=lib.TALLY_EXACT, track_mode=track)
    assert_step_parity(a, b, out, precision)
    assert out[-1][0]["transport"]["segments"] > 2_000_000
    for name in FIELDS_EXACT + FIELDS_TALLIED:
        assert np.array_equal(a.engine.field(name), b.engine.field(name)), name
