import sys
sys.path.insert(0, ".")
import __graft_entry__ as e
from mpimc_b200 import decks, driver, lib
g = lib.ImcLib(e.LIB)
n = 10000000
inputs = decks.suolson(precision="FLOAT16", n_input=n // 2, n_max=n, pairwise=sys.argv[1] if len(sys.argv) > 1 else "FALSE")
sim = driver.setup(inputs, g); sim.save_history = False
for i in range(2):
    r = sim.advance()
print(r["transport"]["kernel_ms"])
