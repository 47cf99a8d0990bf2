import sys
sys.path.insert(0, ".")
import __graft_entry__ as e
from mpimc_b200 import decks, driver, lib
g = lib.ImcLib(e.LIB)
prec, n = sys.argv[1], int(float(sys.argv[2]))
sim = driver.setup(decks.suolson(precision=prec, n_input=n // 5, n_max=n, pairwise="TRUE"), g); sim.save_history = False
for i in range(int(sys.argv[3]) if len(sys.argv) > 3 else 6):
    r = sim.advance()
print(r["transport"]["kernel_ms"], r["transport"]["segments"], sim.engine.num_particles())
