"""BASELINE config 4: Su-Olson 1-D in Float16 / Float32 / Float64 with PAIRWISE tallies at 1e8 particles (NINPUT 2e7),
run to t = 1: relative error of the radiation energy density W(x) against the Float64 run and against the benchmark
table embedded in the reference deck (src/inputs/SuOlson.txt:71-72).

    python scratch/config4_study.py [n_max] [n_input]
"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as e
from mpimc_b200 import decks, driver, lib

n_max = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
n_input = int(float(sys.argv[2])) if len(sys.argv) > 2 else n_max // 5
g = lib.ImcLib(e.LIB)
out = {}
for prec in ("FLOAT64", "FLOAT32", "FLOAT16"):
    sim = driver.setup(decks.suolson(precision=prec, n_input=n_input, n_max=n_max, pairwise="TRUE"), g)
    sim.save_history = False
    t0 = time.time(); seg = 0; modes = set(); kms = 0.0
    while not sim.done():
        r = sim.advance()
        seg += r["transport"]["segments"]; modes.add(r["transport"]["tally_mode"]); kms += r["transport"]["kernel_ms"]
        if abs(float(sim.simvars.t) - 1.0) < 1e-3 or float(sim.simvars.t) > 1.0:
            break
    dt = time.time() - t0
    rad = sim.engine.field("radenergydens").astype(np.float64)
    cent = np.asarray(sim.mesh.centers, dtype=float)
    bench = []
    for x, y in zip(decks.SUOLSON_XBENCH, decks.SUOLSON_YBENCH):
        i = int(np.argmin(np.abs(cent - x)))
        bench.append((x, y, float(rad[max(0, i - 2):i + 3].mean())))
    out[prec] = dict(rad=rad, bench=bench, steps=sim.simvars.step, t=float(sim.simvars.t), wall_s=dt, segments=seg, tally_modes=sorted(modes),
                     tracking_ms_per_step=kms / max(sim.simvars.step, 1), particles=sim.engine.num_particles(),
                     energy_error=r["energy"]["energy_error"])
    print(prec, {k: v for k, v in out[prec].items() if k not in ("rad", "bench")}, flush=True)
ref = out["FLOAT64"]["rad"]
coarse = lambda f: f[:400].reshape(40, 10).mean(axis=1)      # 10-cell bins over the region the wave has reached
res = {"n_max": n_max, "n_input": n_input}
for prec in ("FLOAT32", "FLOAT16"):
    res[prec] = {"rel_l2_vs_f64": float(np.linalg.norm(coarse(out[prec]["rad"]) - coarse(ref)) / np.linalg.norm(coarse(ref)))}
for prec in out:
    b = out[prec]["bench"]
    res.setdefault(prec, {})["max_abs_err_vs_benchmark"] = max(abs(g_ - y) for _, y, g_ in b)
    res[prec]["benchmark_points"] = [(x, y, round(g_, 5)) for x, y, g_ in b]
    res[prec].update({k: out[prec][k] for k in ("steps", "wall_s", "segments", "tally_modes", "tracking_ms_per_step", "particles", "energy_error")})
print(json.dumps(res))
