"""Small end-to-end runs of every kernel family, for compute-sanitizer (memcheck / initcheck / racecheck)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as e
from mpimc_b200 import decks, driver, lib

g = lib.ImcLib(e.LIB)
cases = [
    ("suolson f32 atomic", decks.suolson(precision="FLOAT32", n_input=1500, n_max=20000), dict(tally_mode=lib.TALLY_ATOMIC)),
    ("suolson f16 exact", decks.suolson(precision="FLOAT16", n_input=1500, n_max=20000, pairwise="TRUE"), dict()),
    ("crooked f32 refill atomic", decks.crooked_pipe(precision="FLOAT32", n_input=3000, n_max=60000, cellmin=1, pairwise="FALSE"), dict(track_mode=lib.TRACK_REFILL)),
    ("crooked f64 static fixed", decks.crooked_pipe(precision="FLOAT64", n_input=3000, n_max=60000, cellmin=1, pairwise="FALSE"), dict(track_mode=lib.TRACK_HISTORY, tally_mode=lib.TALLY_FIXED)),
    ("crooked f32 exact", decks.crooked_pipe(precision="FLOAT32", n_input=2000, n_max=60000, cellmin=1, pairwise="TRUE"), dict()),
    ("crooked f32 event", decks.crooked_pipe(precision="FLOAT32", n_input=2000, n_max=60000, cellmin=1, pairwise="FALSE"), dict(track_mode=lib.TRACK_EVENT)),
    ("crooked f32 big mesh global tally", decks.crooked_pipe(precision="FLOAT32", n_input=20000, n_max=200000, cellmin=1, mesh_cells=(160, 160), pairwise="FALSE"), dict()),
    ("marshak rw f32", decks.marshak(precision="FLOAT32", n_cells=64, nonuniform=True, randomwalk="TRUE", n_input=2000, n_max=30000, dx_min=2e-4), dict()),
    ("nonuniform multiscale f64", decks.nonuniform_1d(precision="FLOAT64", n_input=2000), dict()),
    ("marshak rw f32 atomic", decks.marshak(precision="FLOAT32", n_cells=64, nonuniform=True, randomwalk="TRUE", n_input=2000, n_max=30000, dx_min=2e-4, pairwise="FALSE"), dict(tally_mode=lib.TALLY_ATOMIC)),
    ("suolson f32 fixed refill", decks.suolson(precision="FLOAT32", n_input=1500, n_max=20000), dict(tally_mode=lib.TALLY_FIXED, track_mode=lib.TRACK_REFILL)),
    ("nonuniform multiscale f32 atomic", decks.nonuniform_1d(precision="FLOAT32", n_input=2000, pairwise="FALSE"), dict(tally_mode=lib.TALLY_ATOMIC)),
    # round 2: MC_RW under the refill schedule, the block reducer for long sequential sums, the restart point
    ("marshak rw f32 refill fixed", decks.marshak(precision="FLOAT32", n_cells=64, nonuniform=True, randomwalk="TRUE", n_input=2000, n_max=30000, dx_min=2e-4, pairwise="FALSE"), dict(tally_mode=lib.TALLY_FIXED, track_mode=lib.TRACK_REFILL)),
    ("suolson f16 sequential exact (seqblock)", decks.suolson(precision="FLOAT16", n_input=30000, n_max=200000, pairwise="FALSE"), dict()),
    ("crooked f64 refill atomic", decks.crooked_pipe(precision="FLOAT64", n_input=3000, n_max=60000, cellmin=1, mesh_cells=(160, 160), pairwise="FALSE"), dict(track_mode=lib.TRACK_REFILL)),
]
for name, inputs, cfg in cases:
    sim = driver.setup(inputs, g, **cfg)
    sim.save_history = False
    sim.engine.history_enable(2)
    for _ in range(3):
        r = sim.advance()
    sim.engine.checkpoint("save"); sim.advance(); sim.engine.checkpoint("restore"); sim.engine.checkpoint("drop")
    sim.engine.field_native("temp"); sim.engine.particles(); sim.fetch_history()
    print(name, r["transport"]["segments"], r["transport"]["variant"], r["transport"]["tally_mode"], flush=True)
print("done")
