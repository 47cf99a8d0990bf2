"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle on the same seeded inputs.

Per-particle state, event outcomes, segment counts and particle counts must be BIT-EXACT (both sides use
the deterministic math header and the same Philox / tape draws).  Tallied fields are sums whose order
differs in the ATOMIC / FIXED tally modes (atomics vs the reference's sequential / pairwise order), so
they are compared to a tolerance stated per precision: Float64 1e-9, Float32 2e-4, Float16 5e-2 relative
to the field's max; after the comparison the oracle's fields are copied into the engine (imc_set_state)
so that the next step again starts from identical inputs and per-particle parity stays bit-exact.
The EXACT tally mode needs no such synchronisation: see test_exact_tally_mode_*.
"""
import os

import numpy as np
import pytest

from mpimc_b200 import decks, driver, lib

pytestmark = pytest.mark.gpu

TOL = {"FLOAT64": 1e-9, "FLOAT32": 2e-4, "FLOAT16": 5e-2}


def field_close(a, b, tol):
    scale = max(np.max(np.abs(b)), 1e-300)
    return np.max(np.abs(a - b)) <= tol * scale


FIELDS_EXACT = ("fleck", "sigma_a", "sigma_s", "beta", "emittedenergy")
FIELDS_TALLIED = ("energydep", "radenergydens", "matenergydens", "temp", "nrg_inc")


def run_pair(inputs, gpu_lib, oracle_lib, steps, sync=True, precision=None, **cfg):
    a = driver.setup(inputs, gpu_lib, **cfg)
    b = driver.setup(inputs, oracle_lib, **{k: v for k, v in cfg.items() if k not in ("tally_mode", "track_mode")})
    a.save_history = b.save_history = False
    out = []
    for _ in range(steps):
        out.append((a.advance(), b.advance()))
        if sync:
            tol = TOL[precision]
            for name in FIELDS_EXACT:
                assert np.array_equal(a.engine.field(name), b.engine.field(name)), name
            for name in FIELDS_TALLIED:
                fa, fb = a.engine.field(name), b.engine.field(name)
                if precision == "FLOAT16" and name in ("nrg_inc", "matenergydens", "temp"):
                    # differences of Float16 sums: the reference's own accumulation error dominates; bound it by
                    # the size of the terms instead of the (cancelling) result
                    ref = max(np.max(np.abs(b.engine.field("energydep"))), np.max(np.abs(b.engine.field("emittedenergy"))), np.max(np.abs(fb)))
                    assert np.max(np.abs(fa - fb)) <= 4 * tol * ref, (name, np.max(np.abs(fa - fb)), ref)
                    continue
                assert field_close(fa, fb, tol), (name, np.max(np.abs(fa - fb)), np.max(np.abs(fb)))
            a.engine.set_state(temp=b.engine.field("temp"), matenergydens=b.engine.field("matenergydens"),
                               radenergydens=b.engine.field("radenergydens"))
    return a, b, out


def assert_step_parity(a, b, out, precision, check_fields=False):
    for ra, rb in out:
        for key in ("n_new_global", "n_new_local", "n_particles", "n_source"):
            assert ra["source"][key] == rb["source"][key], (key, ra["source"], rb["source"])
        assert ra["source"]["totalenergy"] == rb["source"]["totalenergy"]
        for key in ("segments", "histories", "n_census", "n_absorbed", "n_escaped", "n_rw"):
            assert ra["transport"][key] == rb["transport"][key], (key, ra["transport"], rb["transport"])
    pa, ia = a.engine.particles()
    pb, ib = b.engine.particles()
    assert np.array_equal(ia, ib)
    assert np.array_equal(pa, pb), f"max |diff| = {np.max(np.abs(pa - pb))}"
    if check_fields:
        tol = TOL[precision]
        for name in ("fleck", "sigma_a", "emittedenergy"):
            assert np.array_equal(a.engine.field(name), b.engine.field(name)) or field_close(a.engine.field(name), b.engine.field(name), tol), name
        for name in ("energydep", "radenergydens", "matenergydens", "temp"):
            fa, fb = a.engine.field(name), b.engine.field(name)
            assert field_close(fa, fb, tol), (name, np.max(np.abs(fa - fb)), np.max(np.abs(fb)))


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32", "FLOAT16"])
def test_suolson_steps_bit_exact(gpu_lib, oracle_lib, precision):
    inputs = decks.suolson(precision=precision, n_input=3000, n_max=30000)
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=6, precision=precision)
    assert_step_parity(a, b, out, precision)
    assert out[-1][0]["source"]["n_particles"] > 10000


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32", "FLOAT16"])
def test_crooked_pipe_steps_bit_exact(gpu_lib, oracle_lib, precision):
    es = (1.0,) if precision != "FLOAT16" else (1024.0,)
    inputs = decks.crooked_pipe(precision=precision, n_input=4000, n_max=60000, cellmin=1 if precision == "FLOAT16" else 2, energyscales=es)
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=4, precision=precision)
    assert_step_parity(a, b, out, precision)


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32"])
def test_small_2d_all_boundaries(gpu_lib, oracle_lib, precision):
    for bcs in (("REFLECT", "VACUUM", "REFLECT", "REFLECT"), ("VACUUM", "REFLECT", "VACUUM", "VACUUM"), ("REFLECT",) * 4):
        inputs = decks.small_2d(precision=precision, n_input=2000, bcs=bcs)
        a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=3, precision=precision)
        assert_step_parity(a, b, out, precision)
        esc = sum(r[0]["transport"]["n_escaped"] for r in out)
        assert (esc > 0) == ("VACUUM" in bcs)


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32"])
def test_marshak_and_multiscale(gpu_lib, oracle_lib, precision):
    a, b, out = run_pair(decks.marshak(precision=precision, n_input=3000, n_max=30000), gpu_lib, oracle_lib, steps=5, precision=precision)
    assert_step_parity(a, b, out, precision)
    a, b, out = run_pair(decks.nonuniform_1d(precision=precision, n_input=3000), gpu_lib, oracle_lib, steps=4, precision=precision)
    assert_step_parity(a, b, out, precision)


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32", "FLOAT16"])
def test_random_walk_bit_exact(gpu_lib, oracle_lib, precision):
    """MC_RW (imc_transport.jl:212-479) including its Float64 promotion inside a history (Q3) and the always-kill
    random-walk step (Q1).  The Marshak deck overflows Float16 (sigma_a = 1000/T^3), so Float16 uses the
    infinite-medium deck with RANDOMWALK on."""
    if precision == "FLOAT16":
        inputs = decks.infinite_medium(precision=precision, n_input=3000, n_max=30000, randomwalk="TRUE", energyscales=(1024.0,))
    else:
        inputs = decks.marshak(precision=precision, n_cells=64, nonuniform=True, randomwalk="TRUE", n_input=3000, n_max=30000, dx_min=2e-4)
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=5, precision=precision)
    assert_step_parity(a, b, out, precision)
    assert sum(r[0]["transport"]["n_rw"] for r in out) > 0


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32", "FLOAT16"])
def test_replay_tape_1d(gpu_lib, oracle_lib, precision):
    """Replay mode: both sides consume the same pre-drawn numbers for a fixed particle batch."""
    rng = np.random.default_rng(7)
    T = {"FLOAT64": np.float64, "FLOAT32": np.float32, "FLOAT16": np.float16}[precision]
    inputs = decks.suolson(precision=precision, n_input=1000, n_max=50000)
    n = 5000
    a = driver.setup(inputs, gpu_lib, rng_mode=lib.RNG_TAPE)
    b = driver.setup(inputs, oracle_lib, rng_mode=lib.RNG_TAPE)
    bits = {np.float16: 11, np.float32: 24, np.float64: 53}[T]
    uni = (rng.integers(0, 2 ** bits, size=(48, n)).astype(np.float64)) * 2.0 ** -bits
    exps = rng.exponential(size=(48, n))
    nc = a.mesh.nx
    dx = float(a.mesh.dx[0])
    slots = np.zeros((n, 9))
    slots[:, 0] = slots[:, 2] = rng.integers(1, nc + 1, size=n)
    slots[:, 1] = (rng.random(n) * 0.002).astype(T)
    slots[:, 3] = (rng.random(n) * dx).astype(T)
    mu = (1 - 2 * rng.random(n)).astype(T); mu[mu == 0] = 0.5
    slots[:, 4] = mu; slots[:, 5] = 1.0
    scale = float(np.atleast_1d(a.mesh.energyscales)[0])
    slots[:, 6] = slots[:, 7] = (rng.random(n) * 0.01 * scale + 1e-3 * scale).astype(T); slots[:, 8] = scale
    slots[:50, 2] = 1; slots[:50, 4] = -np.abs(slots[:50, 4])          # reflect at the left wall
    slots[50:100, 2] = nc; slots[50:100, 4] = np.abs(slots[50:100, 4])  # escape through the right wall
    for s in (a, b):
        s.engine.update(0.002)
        s.engine.set_particles(slots)
        s.engine.set_transport_tape(uni, exps)
    ra, rb = a.engine.transport(0.002, 0), b.engine.transport(0.002, 0)
    for key in ("segments", "n_census", "n_absorbed", "n_escaped"):
        assert ra[key] == rb[key]
    eva, nsa = a.engine.outcomes(n)
    evb, nsb = b.engine.outcomes(n)
    assert np.array_equal(eva, evb) and np.array_equal(nsa, nsb)
    assert ra["n_escaped"] > 0
    pa, _ = a.engine.particles(); pb, _ = b.engine.particles()
    alive = pb[:, 7] != -1.0
    assert np.array_equal(pa[alive], pb[alive])
    assert np.array_equal(pa[:, 7] == -1.0, ~alive)
    assert a.engine.clean() == b.engine.clean() == int(alive.sum())
    pa, _ = a.engine.particles(); pb, _ = b.engine.particles()
    assert np.array_equal(pa, pb)


def test_replay_tape_exhaustion_is_an_error(gpu_lib):
    inputs = decks.suolson(precision="FLOAT64", n_input=100, n_max=1000)
    a = driver.setup(inputs, gpu_lib, rng_mode=lib.RNG_TAPE)
    a.engine.update(0.002)
    slots = np.zeros((4, 9)); slots[:, 0] = slots[:, 2] = 5; slots[:, 3] = 0.005; slots[:, 4] = 0.3; slots[:, 5] = 1
    slots[:, 6] = slots[:, 7] = 1.0; slots[:, 8] = 1.0
    a.engine.set_particles(slots)
    a.engine.set_transport_tape(np.full((1, 4), 0.25), np.full((1, 4), 1e-4))
    with pytest.raises(lib.ImcError) as e:
        a.engine.transport(0.002, 0)
    assert e.value.code == -5


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32"])
def test_fixed_point_tally_is_order_free(gpu_lib, oracle_lib, precision):
    """TALLY_FIXED: two runs give bit-identical fields, and they agree with the oracle to the tolerance."""
    inputs = decks.crooked_pipe(precision=precision, n_input=4000, n_max=60000, cellmin=2)
    runs = []
    for _ in range(2):
        a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=3, precision=precision, tally_mode=lib.TALLY_FIXED)
        runs.append({k: a.engine.field(k) for k in ("energydep", "radenergydens", "temp")})
        assert_step_parity(a, b, out, precision)
    for k in runs[0]:
        assert np.array_equal(runs[0][k], runs[1][k]), k


def test_energy_conservation_and_clean_kat(gpu_lib):
    sim = driver.setup(decks.suolson(precision="FLOAT64", n_input=5000, n_max=100000), gpu_lib)
    for _ in range(10):
        r = sim.advance()
        assert abs(r["energy"]["energy_error"]) < 1e-9
    # the reference's clean test (test/runtests.jl:78-87): a particle whose slot 8 is -1.0 is removed
    slots = np.array([[1.0, 2e-4, 3.0, 4e-3, 0.5, 6.0, 7.0, -1.0, 1.0]])
    sim.engine.set_particles(slots)
    assert sim.engine.num_particles() == 1
    assert sim.engine.clean() == 0


@pytest.mark.parametrize("deckname", ["suolson", "crooked_pipe"])
def test_schedules_give_identical_results(gpu_lib, deckname):
    """Static, warp-refill and event-based schedules give bit-identical particles, counters and (fixed-point) tallies."""
    if deckname == "suolson":
        inputs = decks.suolson(precision="FLOAT32", n_input=20000, n_max=200000)
    else:
        inputs = decks.crooked_pipe(precision="FLOAT32", n_input=20000, n_max=200000, cellmin=2)
    modes = (lib.TRACK_HISTORY, lib.TRACK_REFILL, lib.TRACK_EVENT)
    sims = [driver.setup(inputs, gpu_lib, track_mode=m, tally_mode=lib.TALLY_FIXED) for m in modes]
    for _ in range(4):
        recs = [s.advance() for s in sims]
        for r, m in zip(recs, modes):
            assert r["transport"]["variant"] == m
            for key in ("segments", "histories", "n_census", "n_absorbed", "n_escaped"):
                assert r["transport"][key] == recs[0]["transport"][key], (m, key)
    pa, ia = sims[0].engine.particles()
    for s in sims[1:]:   # the event-based schedule appends survivors in a schedule-dependent order only inside a step;
        pb, ib = s.engine.particles()   # the particle arrays themselves stay in list order
        assert np.array_equal(ia, ib) and np.array_equal(pa, pb)
        for name in ("energydep", "radenergydens", "temp"):   # fixed-point tallies: order-free, so identical too
            assert np.array_equal(sims[0].engine.field(name), s.engine.field(name)), name


@pytest.mark.parametrize("pairwise", ["TRUE", "FALSE"])
@pytest.mark.parametrize("case", ["suolson-f32", "suolson-f16", "crooked-f64", "crooked-f32", "nonuniform-f64", "marshak-rw-f32", "infmed-rw-f16"])
def test_exact_tally_mode_is_bit_exact_without_sync(gpu_lib, oracle_lib, case, pairwise):
    """EXACT tally mode: deposits are reduced in the reference's own order (sequential `+=` or Julia's pairwise
    sum), so EVERYTHING — particles, energydep, radenergydens, temperatures — stays bit-identical to the oracle
    over several time steps with no field synchronisation, in all three precisions."""
    deck, prec = case.rsplit("-", 1)
    precision = {"f64": "FLOAT64", "f32": "FLOAT32", "f16": "FLOAT16"}[prec]
    if deck == "suolson":
        inputs = decks.suolson(precision=precision, n_input=3000, n_max=30000, pairwise=pairwise)
    elif deck == "crooked":
        inputs = decks.crooked_pipe(precision=precision, n_input=4000, n_max=60000, cellmin=2, pairwise=pairwise)
    elif deck == "nonuniform":
        inputs = decks.nonuniform_1d(precision=precision, n_input=3000, pairwise=pairwise)
    elif deck == "marshak-rw":
        inputs = decks.marshak(precision=precision, n_cells=64, nonuniform=True, randomwalk="TRUE", n_input=3000, n_max=30000, dx_min=2e-4, pairwise=pairwise)
    else:
        inputs = decks.infinite_medium(precision=precision, n_input=3000, n_max=30000, randomwalk="TRUE", energyscales=(1024.0,), pairwise=pairwise)
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=5, sync=False, tally_mode=lib.TALLY_EXACT)
    assert_step_parity(a, b, out, precision)
    for ra, rb in out:
        assert ra["transport"]["tally_mode"] == lib.TALLY_EXACT
        assert ra["transport"]["lostenergy"] == rb["transport"]["lostenergy"]
        assert ra["tally"] == rb["tally"], (ra["tally"], rb["tally"])
        assert ra["energy"] == rb["energy"]
    for name in FIELDS_EXACT + FIELDS_TALLIED + ("bee",):
        assert np.array_equal(a.engine.field(name), b.engine.field(name)), name


@pytest.mark.parametrize("precision", ["FLOAT16", "FLOAT32", "FLOAT64"])
@pytest.mark.parametrize("deck", ["suolson", "crooked"])
def test_native_precision_field_transfers_on_gpu(gpu_lib, precision, deck):
    """imc_get_field_native / imc_set_state_native (Array{T} straight from / to the device buffers) agree with the
    Float64 calls, before and after mesh.temp turns Float64 (LINEARIZED decks, Q12)."""
    inputs = (decks.suolson(precision=precision, n_input=500, n_max=5000) if deck == "suolson"
              else decks.crooked_pipe(precision=precision, n_input=2000, n_max=20000))
    sim = driver.setup(inputs, gpu_lib)
    eng = sim.engine
    T = lib.PRECISION_DTYPES[{"FLOAT16": lib.F16, "FLOAT32": lib.F32, "FLOAT64": lib.F64}[precision]]
    assert eng.field_dtype("temp") == T
    assert np.array_equal(eng.field_native("temp").astype(np.float64), eng.field("temp").reshape(-1, order="F"))
    sim.advance()
    assert eng.field_dtype("temp") == (np.float64 if deck == "suolson" else T)
    for name in ("temp", "fleck", "sigma_a", "matenergydens", "radenergydens", "energydep", "emittedenergy", "nrg_inc"):
        a, b = eng.field_native(name), eng.field(name).reshape(-1, order="F")
        assert a.dtype == eng.field_dtype(name)
        assert np.array_equal(a.astype(np.float64), b), name
    t, m, r = eng.field_native("temp"), eng.field_native("matenergydens"), eng.field_native("radenergydens")
    eng.set_state_native(temp=t * 2, matenergydens=m * 2, radenergydens=r * 2)
    assert np.array_equal(eng.field_native("temp"), t * 2) and np.array_equal(eng.field_native("matenergydens"), m * 2)
    assert np.array_equal(eng.field("radenergydens").reshape(-1, order="F"), (r * 2).astype(np.float64))
    assert eng.stream() != 0


def test_cached_reciprocal_division_is_ieee_exact(gpu_lib):
    """imc_fastdiv.cuh: the division the 2-D tracking loop performs with a cached reciprocal equals the IEEE quotient
    `a / b` for every operand pair tried on the device (random pairs, tracking-shaped pairs, pairs constructed next to
    rounding boundaries and their neighbours)."""
    import ctypes as C
    f = gpu_lib.dll.imc_cuda_selftest_div
    f.argtypes = [C.c_int, C.c_uint64, C.c_longlong, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(C.c_float)]
    bad, n = C.c_ulonglong(), C.c_ulonglong()
    fb = (C.c_float * 4)()
    assert f(0, 20261017, 100000, C.byref(bad), C.byref(n), fb) == 0
    assert n.value > 1e11 and bad.value == 0, (n.value, bad.value, list(fb))


def test_device_min_fast_paths_match_the_generic_definition(gpu_lib):
    """imc_num.h: Julia's NaN-propagating min (and the NaN-skipping one of imc_transport.jl:551-557) as DMNMX / FMNMX on the
    device against the branchy generic definition, over all pairs of 16 special values (signed zeros, NaNs, infinities,
    subnormals, extremes) and random bit patterns, Float64 and Float32."""
    import ctypes as C
    f = gpu_lib.dll.imc_cuda_selftest_min
    f.argtypes = [C.c_int, C.c_uint64, C.c_longlong, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    bad, n = C.c_ulonglong(), C.c_ulonglong()
    assert f(0, 20261018, 20000, C.byref(bad), C.byref(n)) == 0
    assert n.value > 1e10 and bad.value == 0, (n.value, bad.value)


@pytest.mark.parametrize("deck,precision", [("suolson", "FLOAT16"), ("suolson", "FLOAT64"), ("crooked", "FLOAT32")])
def test_engine_side_history_on_gpu(gpu_lib, deck, precision):
    """imc_history_* on the device: snapshots recorded at the end of every tally equal the per-step downloads."""
    inputs = (decks.suolson(precision=precision, n_input=500, n_max=5000) if deck == "suolson"
              else decks.crooked_pipe(precision=precision, n_input=2000, n_max=20000))
    sim = driver.setup(inputs, gpu_lib)
    sim.save_history = False
    eng = sim.engine
    eng.history_enable(3)
    want = {k: [] for k in ("temp", "matenergydens", "radenergydens", "nrg_inc")}
    for _ in range(5):
        sim.advance()
        for k in want:
            want[k].append(eng.field(k).reshape(-1, order="F").copy())
    assert eng.history_count() == (3, 2)
    for k in want:
        assert np.array_equal(eng.history(k).astype(np.float64), np.array(want[k][:3])), k
    assert np.array_equal(eng.history("radenergydens", first=2, count=1).astype(np.float64), np.array(want["radenergydens"][2:3]))
    assert sim.fetch_history() == 3 and eng.history_count() == (0, 0)


@pytest.mark.parametrize("pairwise", ["TRUE", "FALSE"])
@pytest.mark.parametrize("case", ["suolson-f32", "suolson-f16", "marshak-rw-f32", "nonuniform-f64"])
def test_exact_tally_mode_with_many_records_per_cell(gpu_lib, oracle_lib, case, pairwise):
    """EXACT mode where every cell collects thousands of deposits (the shape of BASELINE config 4): the warp-per-cell
    reduction (coalesced reads, shuffle-fed additions, Julia's pairwise tree above 1024 elements) keeps the reference's
    order of additions — every field bit-identical to the oracle, vacuum losses and Float64 MC_RW deposits included."""
    deck, prec = case.rsplit("-", 1)
    precision = {"f64": "FLOAT64", "f32": "FLOAT32", "f16": "FLOAT16"}[prec]
    if deck == "suolson":
        inputs = decks.suolson(precision=precision, n_input=1_300_000, n_max=4_000_000, pairwise=pairwise)
    elif deck == "marshak-rw":
        inputs = decks.marshak(precision=precision, n_cells=64, nonuniform=True, randomwalk="TRUE", n_input=150_000, n_max=600_000, dx_min=2e-4, pairwise=pairwise)
    else:
        inputs = decks.nonuniform_1d(precision=precision, n_input=150_000, n_max=600_000, pairwise=pairwise)
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=2, sync=False, tally_mode=lib.TALLY_EXACT)
    assert_step_parity(a, b, out, precision)
    for ra, rb in out:
        assert ra["transport"]["tally_mode"] == lib.TALLY_EXACT
        assert ra["transport"]["lostenergy"] == rb["transport"]["lostenergy"]
        assert ra["tally"] == rb["tally"], (ra["tally"], rb["tally"])
        assert ra["energy"] == rb["energy"]
    for name in FIELDS_EXACT + FIELDS_TALLIED + ("bee",):
        assert np.array_equal(a.engine.field(name), b.engine.field(name)), name


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32", "FLOAT16"])
def test_sample_planck_bit_exact(gpu_lib, oracle_lib, precision):
    """Sourcing.sample_planck (imc_sourcing.jl:372-399; dormant in the reference): the device kernel
    against the oracle on Philox draws and on a replay tape, including samples that need many series terms."""
    T = {"FLOAT64": np.float64, "FLOAT32": np.float32, "FLOAT16": np.float16}[precision]
    bits = {np.float16: 11, np.float32: 24, np.float64: 53}[T]
    mk = lambda l, **kw: lib.Engine(lib.Config(precision=lib.PRECISION_IDS[np.dtype(T)], geometry=1, nx=4, seed=2024, **kw), l)
    a, b = mk(gpu_lib), mk(oracle_lib)
    ga, gb = a.sample_planck(200_000, step=5), b.sample_planck(200_000, step=5)
    assert np.array_equal(ga, gb, equal_nan=True)
    assert abs(ga[np.isfinite(ga)].mean() - 3.83223) < 0.05
    rng = np.random.default_rng(3)
    n = 3000
    uni = rng.integers(0, 2 ** bits, size=(5, n)).astype(np.float64) * 2.0 ** -bits
    uni[0, :64] = 1.0 - rng.integers(1, 200, size=64) * 2.0 ** -bits
    uni[3, 100] = 0.0
    a, b = mk(gpu_lib, rng_mode=lib.RNG_TAPE), mk(oracle_lib, rng_mode=lib.RNG_TAPE)
    a.set_source_tape(uni); b.set_source_tape(uni)
    ta, tb = a.sample_planck(n), b.sample_planck(n)
    assert np.array_equal(ta, tb, equal_nan=True) and np.isinf(ta[100])
    a.set_source_tape(uni[:2])
    with pytest.raises(lib.ImcError) as e:
        a.sample_planck(n)
    assert e.value.code == -5


_SKIP_REDUCER_SNIPPET = r"""
import sys
import os

import numpy as np
sys.path.insert(0, "tests")
import __graft_entry__ as entry
from mpimc_b200 import decks, lib
from test_gpu_parity import FIELDS_EXACT, FIELDS_TALLIED, assert_step_parity, run_pair
g, o = lib.ImcLib(entry.LIB), lib.ImcLib(entry.ORACLE_LIB)
cases = [("FLOAT16", decks.suolson(precision="FLOAT16", n_input=60000, n_max=400000, pairwise="FALSE")),
         ("FLOAT16", decks.infinite_medium(precision="FLOAT16", n_input=30000, n_max=200000, pairwise="FALSE", randomwalk="TRUE", energyscales=(1024.0,))),
         ("FLOAT32", decks.suolson(precision="FLOAT32", n_input=60000, n_max=400000, pairwise="FALSE"))]
for precision, d in cases:
    a, b, out = run_pair(d, g, o, steps=4, sync=False, tally_mode=lib.TALLY_EXACT)
    assert_step_parity(a, b, out, precision)
    for name in FIELDS_EXACT + FIELDS_TALLIED:
        assert np.array_equal(a.engine.field(name), b.engine.field(name), equal_nan=True), (precision, name)
print("stagnation-skip reducer: identical")
"""


@pytest.mark.gpu
def test_stagnation_skip_reducer_is_bit_exact(built):
    """warp_seq_add_skip (sequential EXACT sums with their stagnant stretches skipped) against the oracle, in a process
    that has IMC_EXACT_SKIP=1: Float16 Su-Olson (long stagnant chains), a Float16 random-walk deck (Float64 records mixed
    in), and a Float32 deck (hardly any stagnation)."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-c", _SKIP_REDUCER_SNIPPET], env=dict(os.environ, IMC_EXACT_SKIP="1"), capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), timeout=900)
    assert r.returncode == 0 and "identical" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
