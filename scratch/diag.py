import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpimc_b200 import lib, driver, decks
import __graft_entry__ as e
g = lib.ImcLib(e.LIB); o = lib.ImcLib(e.ORACLE_LIB)
prec = sys.argv[1] if len(sys.argv) > 1 else "FLOAT64"
deckname = sys.argv[2] if len(sys.argv) > 2 else "suolson"
if deckname == "suolson": inputs = decks.suolson(precision=prec, n_input=3000, n_max=30000)
elif deckname == "cp": inputs = decks.crooked_pipe(precision=prec, n_input=4000, n_max=60000, cellmin=2)
a = driver.setup(inputs, g); b = driver.setup(inputs, o)
def cmp(tag):
    pa, ia = a.engine.particles(); pb, ib = b.engine.particles()
    if pa.shape != pb.shape: print(tag, "shape", pa.shape, pb.shape); return False
    d = pa != pb
    if d.any():
        rows, cols = np.nonzero(d)
        print(tag, "ndiff", d.sum(), "slots", np.unique(cols), "first rows", rows[:5])
        for r, c in list(zip(rows, cols))[:5]:
            print("   row", r, "slot", c, repr(pa[r, c]), repr(pb[r, c]), "rel", abs(pa[r,c]-pb[r,c])/abs(pb[r,c]), "particle", pb[r])
        return False
    return True
def cmpf(tag, names):
    ok = True
    for n in names:
        fa, fb = a.engine.field(n), b.engine.field(n)
        if not np.array_equal(fa, fb):
            idx = np.argmax(np.abs(fa-fb)); print(tag, n, "differs: max", np.max(np.abs(fa-fb)), "at", idx, fa.flat[idx], fb.flat[idx], "n", (fa!=fb).sum()); ok = False
    return ok
for step in range(6):
    sv = a.simvars
    dt = float(sv.dt)
    for s in (a, b): s.engine.update(dt)
    cmpf(f"step{step} update", ["fleck", "sigma_a", "beta", "bee"])
    ra = a.engine.source(dt, sv.n_input, float(sv.cellmin), step); rb = b.engine.source(dt, sv.n_input, float(sv.cellmin), step)
    if ra != rb: print("source stats differ", ra, rb)
    cmpf(f"step{step} source", ["emittedenergy"])
    cmp(f"step{step} after source")
    ta = a.engine.transport(dt, step); tb = b.engine.transport(dt, step)
    ta.pop("kernel_ms"); tb.pop("kernel_ms"); ta.pop("tally_mode"); tb.pop("tally_mode")
    if ta != tb: print("transport stats differ", ta, tb)
    cmp(f"step{step} after transport")
    a.engine.clean(); b.engine.clean()
    cmp(f"step{step} after clean")
    a.engine.tally(float(sv.t), dt); b.engine.tally(float(sv.t), dt)
    cmpf(f"step{step} tally", ["energydep", "radenergydens", "matenergydens", "temp", "nrg_inc"])
    a.engine.energycheck(); b.engine.energycheck()
    driver.timestep(str(inputs["TIMESTEPPING"]).upper(), a.simvars); driver.timestep(str(inputs["TIMESTEPPING"]).upper(), b.simvars)
print("done")
