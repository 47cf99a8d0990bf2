"""Recorded runs: the fixture format written by tests/golden/record_reference.jl (the REAL reference under Julia) and the
code that replays such a run through an engine (the oracle on CPU, the CUDA engine on the GPU).

A recorded run is a directory:
    deck.txt                      the input deck, in the reference's own format
    manifest.json                 {"producer", "precision", "geometry", "steps"}
    <k>_scalars.f64               [dt, t, totalenergy, lostenergy after transport, totalenergydep] of step k
    <k>_fleck|beta|bee|sigma_a|sigma_s.f64                      after Update.update      (Julia's linear order: x fastest)
    <k>_source.uni.len/.f64       rand(T) draws of each new particle, in call order (ragged: counts + concatenation)
    <k>_after_source.f64          particle list [n, 9|10], every slot as Float64, after Sourcing.sourcing
    <k>_emittedenergy.f64
    <k>_transport.uni|exp.len/.f64   draws of each tracked particle: rand(T) values / randexp values
    <k>_after_transport.f64       particle list after Transport.MC / MC_RW / MC2D (dead particles flagged, not removed)
    <k>_energydep.f64
    <k>_after_clean.f64, <k>_temp|matenergydens|radenergydens.f64   after Clean.clean and Tally.tally

`replay` checks what BASELINE.json's replay mode asks for: particle counts, cell indices, event outcomes (dead flags)
bit-exact; energies, positions, times and directions within `ulps` units in the last place of the deck precision; fields
within `ulps`-scaled relative tolerances (sums of many rounded terms).  With ulps = 0 everything must be identical — that
is what the fixture written by `write_with_refpy` (same deterministic elementary functions as the engine) is held to.
No fixture from real Julia is committed: Julia is not available where this repository is built.
"""
import json
import os
import shutil

import numpy as np

PREC = {"FLOAT16": np.float16, "FLOAT32": np.float32, "FLOAT64": np.float64}
BITS = {np.float16: 11, np.float32: 24, np.float64: 53}


# ------------------------------------------------------------------------------------------------------------ file helpers
def _put(d, k, name, a):
    np.asarray(a, dtype="<f8").ravel().tofile(os.path.join(d, f"{k}_{name}.f64"))


def _put_ragged(d, k, name, rows):
    np.asarray([len(r) for r in rows], dtype="<i8").tofile(os.path.join(d, f"{k}_{name}.len"))
    (np.concatenate([np.asarray(r, dtype="<f8") for r in rows]) if rows else np.zeros(0)).astype("<f8").tofile(os.path.join(d, f"{k}_{name}.f64"))


def _get(d, k, name):
    return np.fromfile(os.path.join(d, f"{k}_{name}.f64"), dtype="<f8")


def _get_ragged(d, k, name, pad, min_rows=1):
    """-> [max draws, n] draw-major tape; unused entries hold `pad` (never consumed when the run matches the record)."""
    lens = np.fromfile(os.path.join(d, f"{k}_{name}.len"), dtype="<i8")
    flat = _get(d, k, name)
    tape = np.full((max(int(lens.max()) if lens.size else 0, min_rows), len(lens)), pad)
    pos = 0
    for j, n in enumerate(lens):
        tape[:n, j] = flat[pos:pos + n]; pos += n
    return tape


# ------------------------------------------------------------------------------------------------------------ writer (Python stand-in for record_reference.jl)
class _RecordingTape:
    """A refpy.Tape that draws from a numpy generator and remembers what it handed out."""

    def __init__(self, rng, T):
        self.rng, self.T, self.u, self.e = rng, T, [], []

    def rand(self, T):
        v = float(self.rng.integers(0, 2 ** BITS[self.T])) * 2.0 ** -BITS[self.T]
        self.u.append(v)
        return T(v)

    def randexp64(self):
        v = float(self.rng.exponential())
        self.e.append(v)
        return np.float64(v)

    def randexp(self, T):
        v = float(T(self.rng.exponential()))          # the record holds the T-valued draw, like Float64(randexp(T)) in Julia
        self.e.append(v)
        return T(v)


def write_with_refpy(out, deck_path, steps, oracle_lib, refpy, seed=0):
    """Writes a recorded run with oracle/imc_refpy.py playing the reference (format check; NOT a Julia fixture)."""
    from mpimc_b200 import deck as _deck
    from mpimc_b200 import driver
    os.makedirs(out, exist_ok=True)
    shutil.copyfile(deck_path, os.path.join(out, "deck.txt"))
    inputs = _deck.read_inputs(deck_path)
    mesh = _deck.mesh_generation(inputs)
    sv = driver.make_simvars(inputs, mesh)
    T = inputs["PRECISION"]
    S = refpy.State(inputs, mesh, sv, _deck.set_constants(inputs), refpy.JuliaMath(oracle_lib.dll))
    rng = np.random.default_rng(seed)
    rw = mesh.geometry == "1D" and str(inputs.get("RANDOMWALK", "FALSE")).upper() == "TRUE"
    if rw:
        refpy.randomwalk_table(S)
    for k in range(steps):
        S.dt, S.t = T(sv.dt), T(sv.t)
        scal = [float(sv.dt), float(sv.t)]
        refpy.update(S)
        for name, f in (("fleck", S.fleck), ("beta", S.beta), ("bee", S.bee), ("sigma_a", S.sigma_a), ("sigma_s", S.sigma_s)):
            _put(out, k, name, S.field(f))
        n0 = len(S.particles)
        made = []

        def tapes():
            while True:
                made.append(_RecordingTape(rng, T))
                yield made[-1]
        refpy.sourcing(S, tapes())
        n_new = len(S.particles) - n0
        _put_ragged(out, k, "source.uni", [t.u for t in made[:n_new]])
        _put(out, k, "after_source", S.slots()); _put(out, k, "emittedenergy", S.field_scaled(S.emittedenergy))
        scal.append(float(S.totalenergy))
        tp = [_RecordingTape(rng, T) for _ in S.particles]
        (refpy.MC_RW if rw else refpy.MC if mesh.geometry == "1D" else refpy.MC2D)(S, tp)
        _put_ragged(out, k, "transport.uni", [t.u for t in tp]); _put_ragged(out, k, "transport.exp", [t.e for t in tp])
        _put(out, k, "after_transport", S.slots()); _put(out, k, "energydep", S.field_scaled(S.energydep))
        scal.append(float(T(S.lostenergy)))
        refpy.clean(S)
        refpy.tally(S)
        _put(out, k, "after_clean", S.slots())
        for name, f in (("temp", S.temp), ("matenergydens", S.matenergydens), ("radenergydens", S.radenergydens)):
            _put(out, k, name, S.field(f))
        scal.append(float(S.totalenergydep))
        _put(out, k, "scalars", scal)
        refpy.energychecker(S)
        driver.timestep(str(inputs["TIMESTEPPING"]).upper(), sv)
    with open(os.path.join(out, "manifest.json"), "w") as f:
        json.dump({"producer": "oracle/imc_refpy.py (format check, not Julia)", "precision": {v: k for k, v in PREC.items()}[T],
                   "geometry": mesh.geometry, "steps": steps}, f)


# ------------------------------------------------------------------------------------------------------------ replay
def _close(got, want, T, ulps, what, scale=1.0):
    got, want = np.asarray(got, dtype=np.float64).ravel(order="F"), np.asarray(want, dtype=np.float64).ravel(order="F")
    assert got.shape == want.shape, f"{what}: {got.shape} vs {want.shape}"
    if ulps == 0:
        ok = (got == want) | (np.isnan(got) & np.isnan(want))
    else:
        tol = ulps * scale * np.abs(np.spacing(want.astype(T)).astype(np.float64))
        ok = (np.abs(got - want) <= tol) | (np.isnan(got) & np.isnan(want)) | (got == want)
    assert ok.all(), f"{what}: {int((~ok).sum())} of {ok.size} differ, first at {int(np.argmin(ok))}: {got[np.argmin(ok)]!r} vs {want[np.argmin(ok)]!r}"


def replay(engine_lib, fixture, ulps=4, field_ulps=None):
    """Drives `engine_lib` through the recorded run.  Returns the number of particle-steps compared."""
    from mpimc_b200 import driver, lib
    man = json.load(open(os.path.join(fixture, "manifest.json")))
    T = PREC[man["precision"]]
    sim = driver.setup(os.path.join(fixture, "deck.txt"), engine_lib, rng_mode=lib.RNG_TAPE, tally_mode=lib.TALLY_EXACT)
    eng = sim.engine
    geom1 = man["geometry"] == "1D"
    nslots = 9 if geom1 else 10
    int_slots = [0, 2] if geom1 else [1, 2]
    dead_slot = 7
    fu = (64 * max(ulps, 0) if field_ulps is None else field_ulps)      # sums of many terms: a looser bound than per-particle values
    compared = 0
    for k in range(man["steps"]):
        dt, t, totalenergy, lost, totaldep = _get(fixture, k, "scalars")[:5]
        eng.update(float(dt))
        for name in ("fleck", "sigma_a", "sigma_s", "beta", "bee"):
            _close(eng.field(name), _get(fixture, k, name), T, ulps, f"step {k} {name}")
        n0 = eng.num_particles()
        eng.set_source_tape(_get_ragged(fixture, k, "source.uni", 0.25, min_rows=1))
        src = eng.source(float(dt), sim.simvars.n_input, float(sim.simvars.cellmin), k)
        want = _get(fixture, k, "after_source").reshape(-1, nslots)
        assert src["n_particles"] == len(want), f"step {k}: {src['n_particles']} particles after sourcing, the record has {len(want)}"
        got = eng.particles()[0]
        assert np.array_equal(got[:, int_slots], want[:, int_slots]), f"step {k}: cell indices after sourcing"
        _close(got, want, T, ulps, f"step {k} particles after sourcing")
        _close([src["totalenergy"]], [totalenergy], T, fu, f"step {k} totalenergy")
        _close(eng.field("emittedenergy"), _get(fixture, k, "emittedenergy"), T, ulps, f"step {k} emittedenergy")
        if ulps:                       # continue from the recorded state, so that differences do not accumulate over stages
            eng.set_particles(want)
        eng.set_transport_tape(_get_ragged(fixture, k, "transport.uni", 0.25), _get_ragged(fixture, k, "transport.exp", 1.0))
        tr = eng.transport(float(dt), k)
        want = _get(fixture, k, "after_transport").reshape(-1, nslots)
        got = eng.particles()[0]
        dead = want[:, dead_slot] == -1.0
        assert np.array_equal(got[:, dead_slot] == -1.0, dead), f"step {k}: census / absorption / escape outcomes"
        assert np.array_equal(got[~dead][:, int_slots], want[~dead][:, int_slots]), f"step {k}: cell indices after transport"
        _close(got[~dead], want[~dead], T, ulps, f"step {k} surviving particles")
        _close([tr["lostenergy"]], [lost], T, fu, f"step {k} lostenergy")
        compared += len(want)
        eng.clean()
        if ulps:
            eng.set_particles(want[~dead])
        eng.tally(float(t), float(dt))
        _close(eng.field("energydep"), _get(fixture, k, "energydep"), T, fu, f"step {k} energydep")
        temp_T = np.float64 if str(sim.inputs["LINEARIZED"]).upper() == "TRUE" else T      # mesh.temp is Float64 after a LINEARIZED tally (Q12)
        for name in ("matenergydens", "temp", "radenergydens"):
            _close(eng.field(name), _get(fixture, k, name), temp_T if name == "temp" else T, fu, f"step {k} {name}")
        eng.energycheck()
        if ulps:                       # restart the next step from the recorded fields
            eng.set_state(temp=_get(fixture, k, "temp"), matenergydens=_get(fixture, k, "matenergydens"), radenergydens=_get(fixture, k, "radenergydens"))
        driver.timestep(str(sim.inputs["TIMESTEPPING"]).upper(), sim.simvars)
        sim.simvars.step += 1
    return compared
