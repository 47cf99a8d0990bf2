"""Small end-to-end runs over the kernels of the second session of round 2 (compaction per 4096-particle block with one barrier,
survivor count, vectorised census tally with the run hand-off, lazy compaction, clean() from the transport's census count), for
compute-sanitizer (memcheck / racecheck):  IMC_LAZY_CLEAN_MIN=0 compute-sanitizer --tool racecheck python scratch/sanitize2.py"""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as e
from mpimc_b200 import decks, driver, lib

g = lib.ImcLib(e.LIB)
cases = [
    ("suolson f32 atomic (lazy: holes pile up)", decks.suolson(precision="FLOAT32", n_input=1501, n_max=20000), dict(tally_mode=lib.TALLY_ATOMIC)),
    ("nonuniform f64 fixed (lazy)", decks.nonuniform_1d(precision="FLOAT64", n_input=1999), dict(tally_mode=lib.TALLY_FIXED)),
    ("small 2-D f32 atomic (compaction every step)", decks.small_2d(precision="FLOAT32", n_input=2001, bcs=("REFLECT", "VACUUM", "REFLECT", "REFLECT")), dict(tally_mode=lib.TALLY_ATOMIC)),
    ("crooked f32 fixed, global accumulators", decks.crooked_pipe(precision="FLOAT32", n_input=9000, n_max=100000, cellmin=1, mesh_cells=(160, 160), pairwise="FALSE"), dict(tally_mode=lib.TALLY_FIXED)),
    ("infinite medium f16 fixed (8-byte vector loads)", decks.infinite_medium(precision="FLOAT16", n_input=3000, n_max=30000, energyscales=(1024.0,)), dict(tally_mode=lib.TALLY_FIXED)),
]
for name, inputs, cfg in cases:
    sim = driver.setup(inputs, g, **cfg)
    sim.save_history = False
    for _ in range(4):
        r = sim.advance()
    n = sim.engine.num_particles()
    sim.engine.checkpoint("save"); sim.advance(); sim.engine.checkpoint("restore"); sim.engine.checkpoint("drop")
    assert sim.engine.num_particles() == n
    p, ids = sim.engine.particles()            # removes the holes first
    assert len(ids) == n
    sim.advance()
    print(name, r["transport"]["segments"], n, flush=True)
print("done")
