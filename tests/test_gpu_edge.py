"""Edge cases of the transport step on the GPU, against the oracle (bit-exact particles, events and counts) or through
size-independent properties at the full BASELINE size:
empty population, degenerate meshes (1 x 1, 1 x N, N x 1), every boundary VACUUM, distance scale != 1 together with
c != 1 (both time divisors), population sizes that are not multiples of the warp / block size, and the 4096 x 4096,
1.25e8-particle crooked pipe (conservation, outcome bookkeeping, schedule independence of the FIXED tallies)."""
import numpy as np
import pytest

from mpimc_b200 import decks, driver, lib
from test_gpu_parity import assert_step_parity, run_pair

pytestmark = pytest.mark.gpu


def degenerate_2d(precision, nx, ny, bcs, n_input=400, distancescale="1.0"):
    f16 = precision == "FLOAT16"   # NMAX must be Float16-representable (Q10); scaled energies keep the weights normal
    d = decks.small_2d(precision=precision, n_input=n_input, n_max=60000 if f16 else 100000, bcs=bcs, energyscales=(1024.0,) if f16 else (1.0,))
    d["XMESHNODES"] = np.linspace(0.0, 1.0, nx + 1).round(10)
    d["YMESHNODES"] = np.linspace(0.0, 2.0, ny + 1).round(10)
    d["DISTANCESCALE"] = distancescale
    return d


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32", "FLOAT16"])
@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (9, 1), (3, 5)])
def test_degenerate_meshes(gpu_lib, oracle_lib, precision, shape):
    for bcs in (("REFLECT",) * 4, ("VACUUM",) * 4, ("VACUUM", "REFLECT", "REFLECT", "VACUUM")):
        inputs = degenerate_2d(precision, shape[0], shape[1], bcs)
        a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=3, precision=precision)
        assert_step_parity(a, b, out, precision)
        esc = sum(r[0]["transport"]["n_escaped"] for r in out)
        assert (esc > 0) == ("VACUUM" in bcs)


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32"])
@pytest.mark.parametrize("distancescale", ["4.0", "3.0"])
def test_distance_scale_and_both_time_divisors(gpu_lib, oracle_lib, precision, distancescale):
    """(dist / ds) / c with ds != 1 and c != 1 (imc_transport.jl:617): both divisions run, one of them by a cached
    reciprocal in Float32."""
    inputs = degenerate_2d(precision, 6, 4, ("REFLECT", "VACUUM", "REFLECT", "REFLECT"), n_input=1500, distancescale=distancescale)
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=3, precision=precision)
    assert_step_parity(a, b, out, precision)
    inputs = decks.suolson(precision=precision, n_input=1500, n_max=20000)
    inputs["DISTANCESCALE"] = distancescale
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=4, precision=precision)
    assert_step_parity(a, b, out, precision)


def test_empty_population(gpu_lib, oracle_lib):
    """transport / clean / tally / energycheck on an engine that holds no particle."""
    for inputs in (decks.suolson(precision="FLOAT32", n_input=100, n_max=1000), decks.small_2d(precision="FLOAT32", n_input=100)):
        for library in (gpu_lib, oracle_lib):
            sim = driver.setup(inputs, library)
            eng = sim.engine
            eng.update(float(sim.simvars.dt))
            eng.set_particles(np.zeros((0, eng.nslots)))
            tr = eng.transport(float(sim.simvars.dt), 0)
            assert tr["segments"] == 0 and tr["histories"] == 0 and tr["n_census"] == 0
            assert eng.clean() == 0 and eng.num_particles() == 0
            st = eng.tally(0.0, float(sim.simvars.dt))
            assert np.all(eng.field("radenergydens") == 0) and st["totalenergydep"] == 0
            slots, ids = eng.particles()
            assert slots.shape == (0, eng.nslots) and ids.shape == (0,)


@pytest.mark.parametrize("n", [1, 31, 33, 255, 257, 1000])
@pytest.mark.parametrize("track", [lib.TRACK_HISTORY, lib.TRACK_REFILL])
def test_ragged_population_sizes(gpu_lib, oracle_lib, n, track):
    """Populations that do not fill a warp / a block, under both history schedules: same particles as the oracle."""
    inputs = decks.crooked_pipe(precision="FLOAT32", n_input=3000, n_max=60000, cellmin=1, pairwise="FALSE")
    a = driver.setup(inputs, gpu_lib, track_mode=track)
    b = driver.setup(inputs, oracle_lib)
    a.advance(); b.advance()
    slots, ids = b.engine.particles()
    keep = np.linspace(0, len(ids) - 1, n).astype(int)
    for s in (a, b):
        s.engine.set_particles(slots[keep], ids[keep])
    ta = a.engine.transport(float(a.simvars.dt), 1)
    tb = b.engine.transport(float(b.simvars.dt), 1)
    for k in ("segments", "histories", "n_census", "n_absorbed", "n_escaped"):
        assert ta[k] == tb[k], (k, ta, tb)
    assert a.engine.clean() == b.engine.clean()
    pa, ia = a.engine.particles(); pb, ib = b.engine.particles()
    assert np.array_equal(ia, ib) and np.array_equal(pa, pb)


def test_full_size_crooked_pipe_properties(gpu_lib):
    """BASELINE config 5 at its per-GPU size (4096 x 4096 cells, NMAX 1.25e8, Float32): properties that do not need the
    oracle — every history ends in exactly one outcome, the census count is the population after clean, the energy
    balance of imc_energycheck.jl:34 closes to Float32 accuracy, and with FIXED tallies the static and the warp-refill
    schedules give bit-identical fields and segment counts."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs 40 GB of device memory")
    inputs = decks.crooked_pipe(precision="FLOAT32", n_input=62_500_000, n_max=125_000_000, cellmin=1, mesh_cells=(4096, 4096), pairwise="FALSE")
    fields = {}
    for track in (lib.TRACK_REFILL, lib.TRACK_HISTORY):
        sim = driver.setup(inputs, gpu_lib, tally_mode=lib.TALLY_FIXED, track_mode=track)
        sim.save_history = False
        segs = []
        for _ in range(2):
            r = sim.advance()
            tr = r["transport"]
            assert tr["histories"] == tr["n_census"] + tr["n_absorbed"] + tr["n_escaped"] and tr["n_errors"] == 0
            assert sim.engine.num_particles() == tr["n_census"]
            assert tr["histories"] > 60_000_000 and tr["segments"] > 10 * tr["histories"]
            assert abs(r["energy"]["energy_error"]) < 2e-3, r["energy"]
            segs.append(tr["segments"])
        fields[track] = (segs, sim.engine.field_native("temp"), sim.engine.field_native("radenergydens"), sim.engine.field_native("energydep"))
        del sim
    a, b = fields[lib.TRACK_REFILL], fields[lib.TRACK_HISTORY]
    assert a[0] == b[0]
    for x, y in zip(a[1:], b[1:]):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("deck", ["suolson", "crooked"])
def test_float16_deck_with_counts_beyond_float16(gpu_lib, oracle_lib, deck):
    """Q10: NINPUT / NMAX above 65504 in a Float16 deck (BASELINE config 4): counts are integers in the engine and the
    per-cell count arithmetic runs in Float32; the engine and the oracle still agree particle for particle."""
    if deck == "suolson":
        inputs = decks.suolson(precision="FLOAT16", n_input=100000, n_max=200000)
    else:
        inputs = decks.crooked_pipe(precision="FLOAT16", n_input=100000, n_max=200000, cellmin=1, energyscales=(1024.0,))
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=3, precision="FLOAT16")
    assert_step_parity(a, b, out, "FLOAT16")
    assert out[-1][0]["source"]["n_particles"] > 65504


@pytest.mark.parametrize("deck", ["suolson", "crooked"])
def test_fused_step_equals_staged_calls(gpu_lib, oracle_lib, deck):
    """imc_step (update -> source -> transport -> clean -> tally -> energycheck in one ABI call) gives exactly the
    particles, fields and statistics of the stage-by-stage calls, on the engine and against the oracle."""
    inputs = (decks.suolson(precision="FLOAT32", n_input=2000, n_max=20000) if deck == "suolson"
              else decks.crooked_pipe(precision="FLOAT32", n_input=3000, n_max=60000, cellmin=1, pairwise="TRUE"))
    sims = []
    for library, fused in ((gpu_lib, True), (gpu_lib, False), (oracle_lib, True)):
        sim = driver.setup(inputs, library, **({"tally_mode": lib.TALLY_EXACT} if library is gpu_lib else {}))
        sim.save_history = False
        sim.fused = fused
        recs = [sim.advance() for _ in range(3)]
        sims.append((sim, recs))
    (a, ra), (b, rb), (c, rc) = sims
    for x, y in ((ra, rb), (ra, rc)):
        for r1, r2 in zip(x, y):
            for stage in ("source", "tally", "energy"):
                assert r1[stage] == r2[stage], (stage, r1[stage], r2[stage])
            for k in ("segments", "histories", "n_census", "n_absorbed", "n_escaped", "lostenergy"):
                assert r1["transport"][k] == r2["transport"][k], k
    for other in (b, c):
        pa, ia = a.engine.particles(); pb, ib = other.engine.particles()
        assert np.array_equal(ia, ib) and np.array_equal(pa, pb)
        for name in ("temp", "matenergydens", "radenergydens", "energydep", "fleck"):
            assert np.array_equal(a.engine.field(name), other.engine.field(name)), name


@pytest.mark.parametrize("precision,nx,ny,tally", [
    ("FLOAT32", 1536, 8, "atomic"),    # Nc = 12288: Float32 accumulators fill 48 KB exactly — the counter slots must be budgeted too
    ("FLOAT32", 1536, 1, "atomic"), ("FLOAT32", 3072, 1, "atomic"), ("FLOAT32", 6144, 1, "atomic"),
    ("FLOAT64", 768, 1, "atomic"), ("FLOAT64", 1536, 1, "atomic"), ("FLOAT64", 3072, 2, "atomic"),
    ("FLOAT32", 768, 8, "fixed"), ("FLOAT32", 3072, 2, "fixed"), ("FLOAT64", 6144, 1, "fixed"),
])
def test_accumulator_sets_at_the_shared_memory_limit(gpu_lib, oracle_lib, precision, nx, ny, tally):
    """Mesh sizes whose shared-memory accumulator sets total exactly 48 KB (or a power-of-two fraction of it): the launch
    must budget the per-thread counter slots inside the limit (round 1 requested 48 KB + 64 B and failed with
    cudaErrorInvalidValue); particles and event counts stay bit-identical to the oracle."""
    inputs = degenerate_2d(precision, nx, ny, ("REFLECT", "VACUUM", "REFLECT", "VACUUM"), n_input=3000)
    mode = {"atomic": lib.TALLY_ATOMIC, "fixed": lib.TALLY_FIXED}[tally]
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=2, precision=precision, tally_mode=mode)
    assert_step_parity(a, b, out, precision)


@pytest.mark.parametrize("deck", ["suolson", "nonuniform_1d", "small_2d", "crooked"])
def test_lazy_compaction_changes_no_result(gpu_lib, oracle_lib, deck, monkeypatch):
    """Clean.clean on large populations leaves a few dead entries in the list instead of copying every survivor
    (imc_engine_impl.cuh, "Lazy compaction"); IMC_LAZY_CLEAN_MIN=0 switches that on for a small deck.  Counts, statistics and
    particles stay bit-identical to the oracle (which always compacts, imc_clean.jl:6-20), the FIXED tallies bit-identical to
    the engine's own always-compacting run — and the lazy run really skipped compactions (fewer kernel launches)."""
    precision = "FLOAT64" if deck == "nonuniform_1d" else "FLOAT32"
    if deck == "suolson":              # three histories end in eight steps: the holes pile up from step to step
        inputs = decks.suolson(precision=precision, n_input=4000, n_max=60000)
    elif deck == "nonuniform_1d":      # 0.05 % of the histories end per step
        inputs = decks.nonuniform_1d(precision=precision, n_input=3000)
    elif deck == "small_2d":           # 10-30 % per step: above the 1/32 threshold, compaction every step
        inputs = decks.small_2d(precision=precision, n_input=3000, bcs=("REFLECT",) * 4)
    else:
        inputs = decks.crooked_pipe(precision=precision, n_input=4000, n_max=60000, cellmin=2)   # most histories end inside the step: compaction every step
    # 1. against the oracle, float tallies (AUTO), every step's fields handed over as in the other parity tests
    monkeypatch.setenv("IMC_LAZY_CLEAN_MIN", "0")
    a, b, out = run_pair(inputs, gpu_lib, oracle_lib, steps=8, precision=precision)
    assert_step_parity(a, b, out, precision)              # exports the particles: the holes are removed first
    died = sum(r[0]["transport"]["n_absorbed"] + r[0]["transport"]["n_escaped"] for r in out)
    assert died > 0
    # 2. lazy against always-compacting on the engine alone, FIXED tallies: every number of the run bit-identical
    runs = {}
    for name, env in (("lazy", "0"), ("eager", str(1 << 62))):
        monkeypatch.setenv("IMC_LAZY_CLEAN_MIN", env)
        sim = driver.setup(inputs, gpu_lib, tally_mode=lib.TALLY_FIXED, track_mode=lib.TRACK_REFILL)   # one tracking launch per step: the launch counts compare
        sim.save_history = False
        recs = [sim.advance() for _ in range(8)]
        launches = sim.engine.kernel_launches()
        runs[name] = (sim, recs, launches, sim.engine.particles())
    (la, lrec, ll, (lp, lid)), (ea, erec, el, (ep, eid)) = runs["lazy"], runs["eager"]
    for r1, r2 in zip(lrec, erec):
        for stage in ("source", "tally", "energy"):
            assert r1[stage] == r2[stage], (stage, r1[stage], r2[stage])
        for k in ("segments", "histories", "n_census", "n_absorbed", "n_escaped", "lostenergy"):
            assert r1["transport"][k] == r2["transport"][k], k
    assert np.array_equal(lid, eid) and np.array_equal(lp, ep)
    for name in ("temp", "matenergydens", "radenergydens", "energydep"):
        assert np.array_equal(la.engine.field(name), ea.engine.field(name)), name
    if deck in ("suolson", "nonuniform_1d"):
        assert ll < el, (ll, el)


@pytest.mark.parametrize("precision", ["FLOAT16", "FLOAT32", "FLOAT64"])
@pytest.mark.parametrize("geom", ["1d", "2d"])
@pytest.mark.parametrize("tally", ["fixed", "atomic"])
def test_census_tally_is_the_per_cell_sum_over_the_particle_list(gpu_lib, precision, geom, tally):
    """Tally.tally's radiation energy density (imc_tally.jl:84-113): radenergydens[cell] = sum over the surviving particles of
    E / (dx [dy] scale).  The kernel reads the list four particles per thread with vector loads and joins runs of equal cells
    across the lanes of a warp; here its result is recomputed from the exported particle list, for every precision (vector
    width 8 / 16 / 2 x 16 bytes), both geometries and both accumulator kinds, with populations that are no multiple of 4."""
    f16 = precision == "FLOAT16"
    es = (1024.0,) if f16 else (1.0,)
    if geom == "1d":
        inputs = decks.nonuniform_1d(precision=precision, n_input=3000) if not f16 else decks.infinite_medium(precision=precision, n_input=3000, n_max=30000, energyscales=es)
    else:
        inputs = decks.small_2d(precision=precision, n_input=3000, n_max=60000 if f16 else 100000, bcs=("REFLECT", "VACUUM", "REFLECT", "REFLECT"), energyscales=es)
    sim = driver.setup(inputs, gpu_lib, tally_mode=lib.TALLY_FIXED if tally == "fixed" else lib.TALLY_ATOMIC)
    sim.save_history = False
    tol = {"FLOAT16": 3e-3, "FLOAT32": 1e-6, "FLOAT64": 1e-13}[precision]   # the terms are rounded to T one by one, the sum once
    if tally == "fixed" and precision == "FLOAT64":
        tol = 1e-8   # every term is rounded to the fixed-point quantum, 2^-62 of a bound on the largest possible sum (FIXED); measured 2e-10
    if tally == "atomic" and precision == "FLOAT32":
        tol = 2e-5   # Float32 shared-memory accumulators, ~100 additions per cell in an order that changes from run to run
    for _ in range(2):
        sim.advance()
        eng = sim.engine
        slots, _ = eng.particles()
        rad = eng.field("radenergydens")
        dx = np.asarray(sim.mesh.dx, dtype=np.float64)
        want = np.zeros(rad.size)
        if geom == "1d":
            cell = slots[:, 2].astype(np.int64) - 1
            np.add.at(want, cell, slots[:, 6] / (dx[cell] * slots[:, 8]))
        else:
            dy = np.asarray(sim.mesh.dy, dtype=np.float64)
            cx, cy = slots[:, 1].astype(np.int64) - 1, slots[:, 2].astype(np.int64) - 1
            np.add.at(want, cx + dx.size * cy, slots[:, 7] / ((dx[cx] * dy[cy]) * slots[:, 9]))
        assert slots.shape[0] > 500 and want.max() > 0
        assert np.max(np.abs(rad.ravel(order="F") - want)) <= tol * want.max(), (np.max(np.abs(rad.ravel(order="F") - want)), want.max())
