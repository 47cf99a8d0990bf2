"""The C++ oracle against a second, independently written restatement of the reference (oracle/imc_refpy.py: plain
Python, numpy scalars standing in for Julia's Float16 / Float32 / Float64 values so that the reference's promotions happen
by themselves).  Both are driven stage by stage — update, sourcing, MC / MC_RW / MC2D, clean, tally, energychecker — on the
same deck, with the same pre-drawn random numbers (replay tapes), over several time steps, and must agree bit for bit in
every particle slot, every field and every scalar.  A misreading of the Julia source would have to be made twice, in two
languages and two sittings, to pass.  (Julia itself cannot run here: SURVEY.md §8c; parity against it stays unpinned.)"""
import importlib.util
import os

import numpy as np
import pytest

import __graft_entry__ as entry
from mpimc_b200 import deck as _deck
from mpimc_b200 import decks, driver, lib

_spec = importlib.util.spec_from_file_location("imc_refpy", os.path.join(entry.ROOT, "oracle", "imc_refpy.py"))
refpy = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(refpy)

BITS = {np.float16: 11, np.float32: 24, np.float64: 53}
N_UNI, N_EXP = 160, 160          # draws available to one particle per stage


def make_tapes(rng, T, n, n_exp=N_EXP):
    uni = rng.integers(0, 2 ** BITS[T], size=(N_UNI, n)).astype(np.float64) * 2.0 ** -BITS[T]
    uni[0, ::17] = 0.0                                       # rand(T) == 0 happens; mu == 0 resampling and mu = sqrt(0) paths
    uni[1, ::29] = 0.5                                       # 1 - 2*0.5 == 0: the isotropic resampling loop
    exps = rng.exponential(size=(n_exp, n))
    return uni, exps


def same(a, b, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:                                   # engine fields come back shaped [nx, (ny), (ns)]: Julia's linear order
        a, b = a.ravel(order="F"), b.ravel(order="F")
    assert a.shape == b.shape, f"{what}: shapes {a.shape} vs {b.shape}"
    if not np.array_equal(a, b, equal_nan=True):
        bad = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {a.size} differ, first at {i}: oracle {a[i]!r} vs refpy {b[i]!r}")


def run_both(inputs, oracle_lib, steps, seed=0, n_batch=0):
    """Runs `steps` time steps on the oracle (through the C ABI) and on imc_refpy, comparing after every stage.
    n_batch > 0: before the first transport call, extra hand-made particles near walls and cell faces are appended."""
    rng = np.random.default_rng(seed)
    sim = driver.setup(inputs, oracle_lib, rng_mode=lib.RNG_TAPE)
    eng, mesh, sv = sim.engine, sim.mesh, sim.simvars
    T = inputs["PRECISION"]
    S = refpy.State(sim.inputs, mesh, sv, _deck.set_constants(sim.inputs), refpy.JuliaMath(oracle_lib.dll))
    geom1 = mesh.geometry == "1D"
    rw = sim.rwvars is not None
    if rw:
        refpy.randomwalk_table(S)
        same(sim.rwvars.aVals, [float(v) for v in S.rw[0]], "aVals")
        same(sim.rwvars.prVals, [float(v) for v in S.rw[1]], "prVals")
        same(sim.rwvars.ptVals, [float(v) for v in S.rw[2]], "ptVals")
    fields = lambda: {k: eng.field(k) for k in ("fleck", "beta", "bee", "sigma_a", "sigma_s")}
    stats = {"events": np.zeros(4, dtype=int), "segments": 0, "sourced": 0}
    for step in range(steps):
        dt, t = float(sv.dt), float(sv.t)
        S.dt, S.t = T(sv.dt), T(sv.t)
        # ---- Update.update
        eng.update(dt); refpy.update(S)
        f = fields()
        for k, d in (("fleck", S.fleck), ("beta", S.beta), ("bee", S.bee), ("sigma_a", S.sigma_a), ("sigma_s", S.sigma_s)):
            same(f[k], S.field(d), f"step {step} update {k}")
        # ---- Sourcing.sourcing
        n_before = eng.num_particles()
        uni, _ = make_tapes(rng, T, 6000, n_exp=1)
        eng.set_source_tape(uni[:16])
        source_tapes = (refpy.Tape(uni[:16, j]) for j in range(uni.shape[1]))
        try:
            src = eng.source(dt, sv.n_input, float(sv.cellmin), step)
        except lib.ImcError as e:
            if e.code != -6:
                raise
            with pytest.raises(refpy.ReferenceThrows):   # NaN / Inf particle count (tointeger) or no representable scale (emittedenergy[i, 0])
                refpy.sourcing(S, source_tapes)
            stats["reference_throws"] = step
            return stats
        refpy.sourcing(S, source_tapes)
        assert src["n_particles"] == len(S.particles), f"step {step}: {src['n_particles']} vs {len(S.particles)} particles after sourcing"
        assert src["totalenergy"] == float(S.totalenergy), f"step {step} totalenergy {src['totalenergy']!r} vs {float(S.totalenergy)!r}"
        same(eng.field("emittedenergy"), S.field_scaled(S.emittedenergy), f"step {step} emittedenergy")
        same(eng.particles()[0], S.slots(), f"step {step} particles after sourcing")
        stats["sourced"] += len(S.particles) - n_before
        stats.setdefault("scales", set()).update(float(p[-1]) for p in S.particles[n_before:])
        if step == 0 and n_batch:
            extra = edge_batch(rng, T, mesh, n_batch)
            slots = np.vstack([S.slots(), extra])
            eng.set_particles(slots)
            S.particles = [[T(v) for v in row] for row in slots]
        # ---- Transport.MC / MC_RW / MC2D
        n = len(S.particles)
        uni, exps = make_tapes(rng, T, n)
        eng.set_transport_tape(uni, exps)
        tr = eng.transport(dt, step)
        tapes = [refpy.Tape(uni[:, j], exps[:, j]) for j in range(n)]
        out = []
        (refpy.MC_RW if rw else refpy.MC if geom1 else refpy.MC2D)(S, tapes, out)
        ev, ns = eng.outcomes(n)
        same(ns, [o[1] for o in out], f"step {step} segments per particle")
        same(ev, [o[0] for o in out], f"step {step} event per particle")
        assert tr["segments"] == sum(o[1] for o in out)
        same(eng.particles()[0], S.slots(), f"step {step} particles after transport (dead ones included)")
        assert tr["lostenergy"] == float(T(S.lostenergy)), f"step {step} lostenergy {tr['lostenergy']!r} vs {float(S.lostenergy)!r}"
        stats["events"] += np.bincount(ev, minlength=4)[:4]; stats["segments"] += tr["segments"]
        # ---- Clean.clean
        alive = eng.clean(); refpy.clean(S)
        assert alive == len(S.particles)
        same(eng.particles()[0], S.slots(), f"step {step} particles after clean")
        # ---- Tally.tally (energydep is compared here: the oracle converts its accumulators in imc_tally)
        tl = eng.tally(t, dt)
        nrg_inc = refpy.tally(S)
        same(eng.field("energydep"), S.field_scaled(S.energydep), f"step {step} energydep")
        same(eng.field("nrg_inc"), S.field(nrg_inc), f"step {step} nrg_inc")
        same(eng.field("matenergydens"), S.field(S.matenergydens), f"step {step} matenergydens")
        same(eng.field("temp"), S.field(S.temp), f"step {step} temp")
        same(eng.field("radenergydens"), S.field(S.radenergydens), f"step {step} radenergydens")
        assert tl["totalenergydep"] == float(S.totalenergydep)
        # ---- EnergyCheck.energychecker
        ec = eng.energycheck()
        rad, err = refpy.energychecker(S)
        assert ec["radenergy"] == float(rad) and (ec["energy_error"] == float(err) or (np.isnan(ec["energy_error"]) and np.isnan(err)))
        driver.timestep(str(sim.inputs["TIMESTEPPING"]).upper(), sv)
        sv.step += 1
    return stats


def edge_batch(rng, T, mesh, n):
    """Particles placed to hit the rarely taken branches: on cell faces (x == 0, x == dx), next to both walls heading out,
    grazing directions, energies at the cut-off, axis-aligned 2-D directions (cos or sin exactly 0 -> Inf / NaN distances)."""
    scale = float(np.atleast_1d(mesh.energyscales)[0])
    E = ((rng.random(n) * 0.01 + 1e-3) * scale).astype(T).astype(np.float64)
    if mesh.geometry == "1D":
        nc = mesh.nx
        s = np.zeros((n, 9))
        s[:, 0] = s[:, 2] = rng.choice([1, 1, nc, nc, max(1, nc // 2)], size=n)
        dx = np.asarray(mesh.dx, dtype=np.float64)[s[:, 2].astype(int) - 1] * float(mesh.distancescale)
        s[:, 1] = (rng.random(n) * float(0.5)).astype(T) * 0.0
        s[:, 3] = (rng.choice([0.0, 1.0, 0.5, 1e-3], size=n) * dx).astype(T)
        s[:, 4] = rng.choice([1.0, -1.0, 0.5, -0.5, 2.0 ** -10, -(2.0 ** -10)], size=n)
        s[:, 5] = 1.0; s[:, 6] = E; s[:, 7] = E; s[:, 8] = scale
        s[::7, 6] = (s[::7, 7] * 0.0101).astype(T)           # just above the 1 % cut-off
        return s
    nx, ny = mesh.nx, mesh.ny
    s = np.zeros((n, 10))
    s[:, 1] = rng.choice([1, nx], size=n); s[:, 2] = rng.choice([1, ny, max(1, ny // 2)], size=n)
    dx = np.asarray(mesh.dx, dtype=np.float64)[s[:, 1].astype(int) - 1] * float(mesh.distancescale)
    dy = np.asarray(mesh.dy, dtype=np.float64)[s[:, 2].astype(int) - 1] * float(mesh.distancescale)
    s[:, 3] = (rng.choice([0.0, 1.0, 0.5], size=n) * dx).astype(T); s[:, 4] = (rng.choice([0.0, 1.0, 0.25], size=n) * dy).astype(T)
    ang = rng.choice([0.0, np.pi / 2, np.pi, -np.pi / 2, np.pi / 4, 3 * np.pi / 4, -3 * np.pi / 4, 1e-3, 3.0], size=n)
    # sin(0.0) == 0 exactly: with y == 0 as well the y distance is NaN, the reference takes its y branch (Q15) without ever
    # moving, and loops for ever at a REFLECT wall — a property of the reference, so that combination is left out
    on_face = (ang == 0.0) & (s[:, 4] == 0.0)
    s[on_face, 4] = (0.25 * dy[on_face]).astype(T)
    s[:, 5] = ang.astype(T); s[:, 6] = 1.0; s[:, 7] = E; s[:, 8] = E; s[:, 9] = scale
    s[::7, 7] = (s[::7, 8] * 0.0101).astype(T)
    return s


F16S = (1024.0,)


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32", "FLOAT16"])
@pytest.mark.parametrize("pairwise", ["FALSE", "TRUE"])
def test_suolson_1d(oracle_lib, precision, pairwise):
    """MC on the LINEARIZED Su-Olson deck: radiation source, REFLECT left wall, temp turning Float64 (Q12)."""
    kw = dict(energyscales=(32768.0,)) if precision == "FLOAT16" else {}
    d = decks.suolson(precision=precision, n_input=150, n_max=2000, pairwise=pairwise, **kw)
    st = run_both(d, oracle_lib, steps=4, seed=1, n_batch=60)
    assert st["events"][0] > 0 and st["events"][1] > 0 and st["sourced"] > 300


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32", "FLOAT16"])
def test_infinite_medium_vacuum_1d(oracle_lib, precision):
    """MC with collisions, cut-offs and escapes through a VACUUM wall; NMAX cap reached (Q9)."""
    d = decks.infinite_medium(precision=precision, n_input=120, n_max=400, pairwise="FALSE", energyscales=(1.0,) if precision != "FLOAT16" else F16S)
    d["SIGMA_A_VALS"] = ["100.0"]; d["RIGHTBC"] = "VACUUM"
    st = run_both(d, oracle_lib, steps=5, seed=2, n_batch=60)
    assert st["events"][2] > 0 and st["events"][1] > 0


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32"])
def test_marshak_surface_source_multiscale(oracle_lib, precision):
    """Surface sources (cosine law, Q30 / Q32), the MARSHAK WAVE opacity quirk (Q18), RAMP time steps, non-uniform cells."""
    st = run_both(decks.marshak(precision=precision, n_cells=24, nonuniform=True, n_input=150, n_max=3000, dx_min=2e-4), oracle_lib, steps=4, seed=3, n_batch=40)
    assert st["sourced"] > 300
    st = run_both(decks.nonuniform_1d(precision=precision, n_input=150, n_max=3000), oracle_lib, steps=3, seed=4)
    assert st["sourced"] > 200


@pytest.mark.parametrize("pairwise", ["FALSE", "TRUE"])
def test_float16_scale_selection(oracle_lib, pairwise):
    """Float16 with several ENERGYSCALES: sorter takes the largest scale whose product stays finite (Q29), so particles of
    different scales coexist and every tally has one plane per scale (Q28) — the regime the scales exist for."""
    d = decks.nonuniform_1d(precision="FLOAT16", n_input=150, n_max=3000, pairwise=pairwise)       # nine scales, 32768 ... 0.5
    st = run_both(d, oracle_lib, steps=4, seed=21, n_batch=40)
    assert len(st["scales"]) >= 2 and st["sourced"] > 300
    d = decks.suolson(precision="FLOAT16", n_input=150, n_max=2000, pairwise=pairwise, energyscales=(32768.0, 1024.0, 32.0, 1.0))
    st = run_both(d, oracle_lib, steps=4, seed=22, n_batch=40)
    assert st["events"][0] > 0 and st["events"][1] > 0


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32", "FLOAT16"])
def test_random_walk_1d(oracle_lib, precision):
    """MC_RW: Float64 promotion inside a history (Q3), deposits not divided by dx (Q2), the always-kill step (Q1)."""
    if precision == "FLOAT16":
        d = decks.infinite_medium(precision=precision, n_input=150, n_max=3000, randomwalk="TRUE", energyscales=F16S)
    else:
        d = decks.marshak(precision=precision, n_cells=24, nonuniform=True, randomwalk="TRUE", n_input=150, n_max=3000, dx_min=2e-4)
    st = run_both(d, oracle_lib, steps=4, seed=5, n_batch=40)
    assert st["events"][3] > 0, "no random-walk step was taken"


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32", "FLOAT16"])
@pytest.mark.parametrize("pairwise", ["FALSE", "TRUE"])
def test_small_2d_all_walls(oracle_lib, precision, pairwise):
    """MC2D with surface sources on all four sides (Q7, Q32), REFLECT and VACUUM walls, axis-aligned directions (Q15)."""
    d = decks.small_2d(precision=precision, n_input=150, n_max=3000, bcs=("REFLECT", "VACUUM", "VACUUM", "REFLECT"), pairwise=pairwise,
                       energyscales=(1.0,) if precision != "FLOAT16" else F16S)
    st = run_both(d, oracle_lib, steps=3, seed=6, n_batch=80)
    assert st["events"][2] > 0 and st["events"][0] > 0


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32"])
def test_crooked_pipe_2d(oracle_lib, precision):
    """The shipped crooked-pipe geometry (thick and thin regions, graded mesh) at a small particle count."""
    st = run_both(decks.crooked_pipe(precision=precision, n_input=1500, n_max=20000, cellmin=0), oracle_lib, steps=4, seed=7)
    assert st["segments"] > 1000


@pytest.mark.parametrize("precision", ["FLOAT64", "FLOAT32"])
@pytest.mark.parametrize("distancescale", ["3.0", "4.0"])
def test_distance_scale(oracle_lib, precision, distancescale):
    """DISTANCESCALE != 1: cell widths scaled inside the loop, `(dist / ds) / c` (imc_transport.jl:125, :611), opacities
    divided by ds at mesh generation, ds as a factor of the sorter products; a several-scale ENERGYSCALES list (Q28 / Q29)."""
    d = decks.suolson(precision=precision, n_input=150, n_max=2000)
    d["DISTANCESCALE"] = distancescale
    d["PHYS_C"] = "2.7"
    run_both(d, oracle_lib, steps=3, seed=8, n_batch=40)
    d = decks.small_2d(precision=precision, n_input=150, n_max=3000, bcs=("VACUUM", "REFLECT", "REFLECT", "VACUUM"))
    d["DISTANCESCALE"] = distancescale
    d["PHYS_C"] = "2.7"
    st = run_both(d, oracle_lib, steps=3, seed=9, n_batch=60)
    assert st["events"][2] > 0


@pytest.mark.parametrize("seed", range(36))
def test_random_decks(oracle_lib, seed):
    """The randomised decks of tests/test_gpu_fuzz.py (mesh shape and grading, boundary conditions, opacities and powers,
    scattering, sources, energy scales, distance scale, c, dt, precision, PAIRWISE) at a small particle count."""
    import test_gpu_fuzz as fuzz
    global N_UNI, N_EXP
    rng = np.random.default_rng(5000 + seed)
    precision = ["FLOAT64", "FLOAT32", "FLOAT16"][seed % 3]
    d = fuzz.random_2d(rng, precision) if seed % 4 != 3 else fuzz.random_1d(rng, precision)
    d["NINPUT"] = str(int(rng.integers(40, 160)))
    if d["GEOMETRY"] == "2D" and len(d["YMESHNODES"]) > len(d["XMESHNODES"]):
        # left / right surface sources read mesh.dx[j] for j up to Ny (Q7): with Ny > Nx the reference throws BoundsError
        # (imc_refpy raises it too; the engine and the C++ oracle read dy[j] there instead) — keep Ny <= Nx
        d["XMESHNODES"], d["YMESHNODES"] = d["YMESHNODES"], d["XMESHNODES"]
        (x1, x2, y1, y2), = d["T_SURFACE_REGS"]
        d["T_SURFACE_REGS"] = [(y1, y2, x1, x2)]
    if d["GEOMETRY"] == "1D":
        d["CELLMIN"] = "0"                               # the 1-D decks have 100-1000 cells: no particle floor per cell
    keep = N_UNI, N_EXP
    N_UNI = N_EXP = 1536                                 # scattering-dominated decks make long histories
    try:
        st = run_both(d, oracle_lib, steps=2, seed=seed, n_batch=24)
    except lib.ImcError as e:
        if e.code != -5:
            raise
        pytest.skip("a history needs more draws than the tape holds")
    finally:
        N_UNI, N_EXP = keep
    assert st["segments"] > 0 or "reference_throws" in st


# ---------------------------------------------------------------------------------------------- Sourcing.sample_planck
def _planck_engine(oracle_lib, precision, **kw):
    return lib.Engine(lib.Config(precision=lib.PRECISION_IDS[np.dtype(precision)], geometry=1, nx=4, seed=77, **kw), oracle_lib)


@pytest.mark.parametrize("T", [np.float64, np.float32, np.float16])
def test_sample_planck_on_a_tape(oracle_lib, T):
    """Sourcing.sample_planck (imc_sourcing.jl:372-399), unused by the reference's step: both restatements on pre-drawn
    numbers, including first draws so close to 1 that the series runs for many terms, zeros (log(0) -> Inf) and, in
    Float32, a first draw the series can never reach (the reference would loop for ever: NaN by convention)."""
    rng = np.random.default_rng(11)
    n = 400
    uni = rng.integers(0, 2 ** BITS[T], size=(5, n)).astype(np.float64) * 2.0 ** -BITS[T]
    uni[0, :40] = 1.0 - rng.integers(1, 60, size=40) * 2.0 ** -BITS[T]      # largest rand(T) values
    uni[2, 50] = 0.0
    if T is np.float32:
        uni[0, 0] = 1.0 - 2.0 ** -24
    eng = _planck_engine(oracle_lib, T, rng_mode=lib.RNG_TAPE)
    eng.set_source_tape(uni)
    got = eng.sample_planck(n)
    m = refpy.JuliaMath(oracle_lib.dll)
    want = np.array([float(refpy.sample_planck(T, refpy.Tape(uni[:, i]), m)) for i in range(n)])
    same(got, want, "sample_planck")
    assert np.isinf(got[50]) and not np.isnan(got[40:]).any()       # Float16 products of four draws underflow to 0 now and then: Inf
    if T is np.float32:
        assert np.isnan(got[0])                                          # 1 - 2^-24 > 0.9999989 = the largest 90 nsum / pi^4 in Float32
    with pytest.raises(lib.ImcError) as e:                                # the accepting branch needs four more draws
        eng.set_source_tape(uni[:3]); eng.sample_planck(n)
    assert e.value.code == -5


@pytest.mark.parametrize("T", [np.float64, np.float32, np.float16])
def test_sample_planck_philox_and_spectrum(oracle_lib, T):
    """Philox draws (stream 3 of sample i): the restatements agree, and the sample mean is the Planck mean
    360 zeta(5) / pi^4 = 3.83223 (Fleck-Cummings series method)."""
    eng = _planck_engine(oracle_lib, T)
    n = 40000
    got = eng.sample_planck(n, step=3)
    f = oracle_lib.dll.imc_oracle_draws
    f.restype = None
    f.argtypes = [refpy.C.c_int32, refpy.C.c_int64, refpy.C.c_uint64, refpy.C.c_uint32, refpy.C.c_uint32, refpy.C.c_int32, refpy._DP, refpy.C.c_int64]
    m = refpy.JuliaMath(oracle_lib.dll)
    draws = np.empty(64)
    for i in range(300):
        f(lib.PRECISION_IDS[np.dtype(T)], 77, i, 3, 3, 0, draws.ctypes.data_as(refpy._DP), 64)
        assert got[i] == float(refpy.sample_planck(T, refpy.Tape(draws), m, max_terms=50)) or np.isnan(got[i])
    ok = np.isfinite(got)
    assert ok.mean() > (0.999 if T is not np.float16 else 0.99)       # Float16: a product of four 11-bit draws underflows to 0 (-> Inf) in ~0.2 %
    assert abs(got[ok].mean() - 3.83223) < (0.05 if T is not np.float16 else 0.08)


REF_INPUTS = "/root/reference/src/inputs"


@pytest.mark.skipif(not os.path.isdir(REF_INPUTS), reason="reference decks not present (GPU box)")
@pytest.mark.parametrize("fname,ninput", [("SuOlson.txt", "1000"), ("InfiniteMedium.txt", "400"), ("MarshakWave.txt", "400"), ("1DNonUniform.txt", "400"),
                                          ("2DNonUniform.txt", "1500"), ("CrookedPipe.txt", "1500")])
def test_shipped_decks(oracle_lib, fname, ninput):
    """The reference's own deck files (src/inputs/*.txt: parser, mesh generator and every stage) with the particle count
    reduced to what Python loops can track; SuOlson.txt runs exactly as shipped (FLOAT16, NINPUT 1000).  Lattice.txt cannot
    run in the reference either (Q21)."""
    d = _deck.read_inputs(os.path.join(REF_INPUTS, fname))
    d["NINPUT"] = ninput
    if fname != "SuOlson.txt":
        d["CELLMIN"] = "0" if fname in ("CrookedPipe.txt", "2DNonUniform.txt") else d["CELLMIN"]   # thousands of cells x CELLMIN otherwise
    global N_UNI, N_EXP
    keep = N_UNI, N_EXP
    N_UNI = N_EXP = 1024
    try:
        st = run_both(d, oracle_lib, steps=3, seed=31)
    finally:
        N_UNI, N_EXP = keep
    assert st["segments"] > 500
