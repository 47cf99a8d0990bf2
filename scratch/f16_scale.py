import sys
sys.path.insert(0, ".")
import __graft_entry__ as e
from mpimc_b200 import decks, driver, lib
g = lib.ImcLib(e.LIB)
for prec in ("FLOAT16", "FLOAT32"):
    for n in (100000, 1000000, 10000000):
        inputs = decks.suolson(precision=prec, n_input=n // 2, n_max=n, pairwise="FALSE")
        sim = driver.setup(inputs, g); sim.save_history = False
        for i in range(3):
            r = sim.advance()
        t = r["transport"]
        print(prec, n, {k: t[k] for k in ("segments", "histories", "n_census", "n_absorbed", "n_errors", "kernel_ms", "variant", "tally_mode")}, "emit", r["source"]["totalenergy"], flush=True)
