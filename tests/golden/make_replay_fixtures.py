"""Generates the replay fixtures in this directory from the CPU oracle.

    python tests/golden/make_replay_fixtures.py

Each fixture is one fixed particle batch plus pre-drawn random numbers (the tape) and the outcome the
oracle computes for it: per-particle event codes, segment counts, final slots, and the energy-deposition
field.  Tests then check (a) on CPU that the oracle still reproduces them and (b) on the GPU that the CUDA
engine, fed the same batch and tape through the C ABI, gives bit-identical particles / events (replay mode,
BASELINE.json north_star).  There are no golden vectors in the reference's own tests for this path and
Julia cannot run here; if a Julia-equipped machine becomes available, the same file format
can be filled by recording rand/randexp around Transport.MC / MC2D.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402
from mpimc_b200 import decks, driver, lib  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DT = {"FLOAT64": np.float64, "FLOAT32": np.float32, "FLOAT16": np.float16}
BITS = {"FLOAT64": 53, "FLOAT32": 24, "FLOAT16": 11}


def deck_for(kind, precision):
    if kind == "1d":  # infinite-medium deck with sigma_a = 100 and a vacuum right wall: collisions, cut-offs, escapes, census
        d = decks.infinite_medium(precision=precision, n_input=100, n_max=10000, energyscales=(1.0,) if precision != "FLOAT16" else (1024.0,))
        d["SIGMA_A_VALS"] = ["100.0"]; d["RIGHTBC"] = "VACUUM"
        return d, 0.0005
    return decks.small_2d(precision=precision, n_input=100, n_max=10000, bcs=("REFLECT", "VACUUM", "VACUUM", "REFLECT"),
                          energyscales=(1.0,) if precision != "FLOAT16" else (1024.0,)), 0.01


def make_batch(kind, precision, sim, n, rng):
    T = DT[precision]
    scale = float(np.atleast_1d(sim.mesh.energyscales)[0])
    if kind == "1d":
        nc, dx = sim.mesh.nx, float(sim.mesh.dx[0])
        s = np.zeros((n, 9))
        s[:, 0] = s[:, 2] = rng.integers(1, nc + 1, size=n)
        s[:, 1] = (rng.random(n) * 0.0005).astype(T)
        s[:, 3] = (rng.random(n) * dx).astype(T)
        mu = (1 - 2 * rng.random(n)).astype(T); mu[mu == 0] = 0.5
        s[:, 4] = mu; s[:, 5] = 1.0
        s[:, 6] = s[:, 7] = ((rng.random(n) * 0.01 + 1e-3) * scale).astype(T); s[:, 8] = scale
        s[: n // 8, 2] = 1; s[: n // 8, 4] = -np.abs(s[: n // 8, 4])
        s[n // 8: n // 4, 2] = nc; s[n // 8: n // 4, 4] = np.abs(s[n // 8: n // 4, 4])
        return s
    nx, ny = sim.mesh.nx, sim.mesh.ny
    s = np.zeros((n, 10))
    s[:, 0] = (rng.random(n) * 0.01).astype(T)
    s[:, 1] = rng.integers(1, nx + 1, size=n); s[:, 2] = rng.integers(1, ny + 1, size=n)
    dx = np.asarray(sim.mesh.dx, dtype=float)[s[:, 1].astype(int) - 1]; dy = np.asarray(sim.mesh.dy, dtype=float)[s[:, 2].astype(int) - 1]
    s[:, 3] = (rng.random(n) * dx).astype(T); s[:, 4] = (rng.random(n) * dy).astype(T)
    s[:, 5] = (2 * np.pi * rng.random(n)).astype(T); s[:, 6] = 1.0
    s[:, 7] = s[:, 8] = ((rng.random(n) * 0.01 + 1e-3) * scale).astype(T); s[:, 9] = scale
    return s


def run(engine_lib, kind, precision, slots, uni, exps):
    inputs, dt = deck_for(kind, precision)
    sim = driver.setup(inputs, engine_lib, rng_mode=lib.RNG_TAPE, tally_mode=lib.TALLY_EXACT)
    sim.engine.update(dt)
    sim.engine.set_particles(slots)
    sim.engine.set_transport_tape(uni, exps)
    st = sim.engine.transport(dt, 0)
    ev, ns = sim.engine.outcomes(len(slots))
    out, _ = sim.engine.particles()
    return dict(events=ev, nseg=ns, slots_out=out, segments=st["segments"], lostenergy=st["lostenergy"], sim=sim)


def main():
    olib = lib.ImcLib(entry.build_oracle())
    n, depth = 256, 40
    for kind in ("1d", "2d"):
        for precision in ("FLOAT64", "FLOAT32", "FLOAT16"):
            rng = np.random.default_rng(hash((kind, precision)) % 2 ** 32 if False else {"1d": 11, "2d": 23}[kind] + BITS[precision])
            inputs, dt = deck_for(kind, precision)
            sim = driver.setup(inputs, olib, rng_mode=lib.RNG_TAPE)
            slots = make_batch(kind, precision, sim, n, rng)
            uni = rng.integers(0, 2 ** BITS[precision], size=(depth, n)).astype(np.float64) * 2.0 ** -BITS[precision]
            exps = rng.exponential(size=(depth, n)) * rng.choice([1.0, 0.05, 8.0], size=(depth, n))
            r = run(olib, kind, precision, slots, uni, exps)
            # the deposit field is only defined after tally; take it straight from the transport accumulators
            r["sim"].engine.clean(); r["sim"].engine.tally(0.0, dt)
            energydep = r["sim"].engine.field("energydep")
            path = os.path.join(HERE, f"replay_{kind}_{precision.lower()}.npz")
            np.savez_compressed(path, kind=kind, precision=precision, slots_in=slots, uniforms=uni, exponentials=exps,
                                events=r["events"], nseg=r["nseg"], slots_out=r["slots_out"], segments=r["segments"],
                                lostenergy=r["lostenergy"], energydep=energydep)
            ev = r["events"]
            print(f"{os.path.basename(path)}: {n} particles, {int(r['segments'])} segments, census {np.sum(ev == 0)}, "
                  f"absorbed {np.sum(ev == 1)}, escaped {np.sum(ev == 2)}, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
