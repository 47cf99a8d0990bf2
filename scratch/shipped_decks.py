"""The shipped decks at their shipped sizes (deck generators reproduce src/inputs/*.txt): wall time per step on the
engine (AUTO modes: EXACT tallies for PAIRWISE / Float16 decks) and on the single-threaded oracle."""
import sys, time
sys.path.insert(0, ".")
import __graft_entry__ as e
from mpimc_b200 import decks, driver, lib
g, o = lib.ImcLib(e.LIB), lib.ImcLib(e.ORACLE_LIB)
cases = [("SuOlson.txt (Float16, 1000 cells)", decks.suolson("FLOAT16", n_input=1000, n_max=50000, pairwise="TRUE"), 60),
         ("CrookedPipe.txt (Float64, 106 x 47)", decks.crooked_pipe(), 12),
         ("MarshakWave.txt (Float64, 300 cells)", decks.marshak("FLOAT64"), 60),
         ("InfiniteMedium.txt (Float32)", decks.infinite_medium(), 40)]
for name, inputs, steps in cases:
    row = [name]
    for tag, library in (("engine", g), ("oracle", o)):
        sim = driver.setup(inputs, library); sim.save_history = False
        for _ in range(3): sim.advance()
        t0 = time.perf_counter(); seg = 0
        for _ in range(steps):
            r = sim.advance(); seg += r["transport"]["segments"]
        dt = time.perf_counter() - t0
        row.append(f"{tag}: {1e3 * dt / steps:.2f} ms/step, {seg / dt:.3g} seg/s ({seg / steps:.3g} seg/step, mode {r['transport']['tally_mode']})")
    print(" | ".join(row), flush=True)
