"""Generates tests/golden/run_digests.json from the CPU oracle: SHA-256 digests of the particle list and of the tallied
fields after a few whole time steps (update -> source -> transport -> clean -> tally -> energycheck, Philox draws) of small
decks in the three precisions.

    python tests/golden/make_run_digests.py

The replay fixtures pin the tracking loop for given draws; these digests pin everything else as well — sourcing, the
Philox keying, the word -> uniform / exponential conversions, the shared elementary functions, the Julia-order sums.
The oracle and the CUDA engine share csrc/imc_math.h and csrc/imc_rng.h, so a change there moves both and the engine-vs-
oracle parity tests stay green; these digests make such a change visible (regenerate them deliberately, and say why in the
commit).  CPU: the oracle must reproduce them.  GPU: the engine (EXACT tallies) must reproduce the same digests."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402
from mpimc_b200 import decks, driver, lib  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "run_digests.json")
STEPS = 4
FIELDS = ("temp", "matenergydens", "radenergydens", "energydep", "emittedenergy", "fleck", "sigma_a")


def cases():
    for precision in ("FLOAT64", "FLOAT32", "FLOAT16"):
        f16 = precision == "FLOAT16"
        yield f"suolson-{precision}", decks.suolson(precision=precision, n_input=800, n_max=6000, pairwise="TRUE")
        yield f"crooked-{precision}", decks.crooked_pipe(precision=precision, n_input=1500, n_max=30000, cellmin=1, pairwise="FALSE",
                                                         **({"energyscales": (1024.0,)} if f16 else {}))
        yield f"small2d-{precision}", decks.small_2d(precision=precision, n_input=600, n_max=30000, bcs=("REFLECT", "VACUUM", "VACUUM", "REFLECT"),
                                                     pairwise="TRUE", energyscales=(1024.0,) if f16 else (4.0, 1.0, 0.5))
        if not f16:
            yield f"marshak-rw-{precision}", decks.marshak(precision=precision, n_cells=64, nonuniform=True, randomwalk="TRUE", n_input=800,
                                                           n_max=20000, dx_min=2e-4, pairwise="TRUE")
            yield f"nonuniform-{precision}", decks.nonuniform_1d(precision=precision, n_input=800, pairwise="FALSE")


def digest(library, inputs, **cfg):
    sim = driver.setup(inputs, library, **cfg)
    sim.save_history = False
    segs = []
    for _ in range(STEPS):
        r = sim.advance()
        segs.append(int(r["transport"]["segments"]))
    slots, ids = sim.engine.particles()
    h = {"segments": segs, "n_particles": int(len(ids)),
         "particles": hashlib.sha256(np.ascontiguousarray(slots).tobytes() + np.ascontiguousarray(ids).tobytes()).hexdigest()}
    for name in FIELDS:
        h[name] = hashlib.sha256(np.ascontiguousarray(sim.engine.field(name), dtype=np.float64).tobytes()).hexdigest()
    return h


def main():
    olib = lib.ImcLib(entry.build_oracle())
    out = {name: digest(olib, inputs) for name, inputs in cases()}
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print(f"wrote {len(out)} digests to {OUT}")


if __name__ == "__main__":
    main()
