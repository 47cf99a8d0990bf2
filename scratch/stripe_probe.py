"""One GPU playing rank 0 of a W-GPU weak-scaling run of crookedpipe_f32 (the other ranks are taken to be statistically
identical: the all-reduces are replaced by a multiplication by W), to see what the sharding of the new particles does to the
tracking kernel:  python scratch/stripe_probe.py W [steps]   (IMC_STRIPE = ordinals per stripe)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as entry
from mpimc_b200 import driver, lib, dist as imc_dist

W = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
w = bench.WORKLOADS["crookedpipe_f32"]
inputs = bench.make_inputs(w, w["particles"] * W, w["mesh"], W, "FALSE", "weak")
glib = lib.ImcLib(entry.LIB)
sim = driver.setup(inputs, glib, device=0, rank=0, world=W)
sim.save_history = False
eng = sim.engine
out = []
for s in range(steps):
    inp, mesh, sv, parts = sim.inputs, sim.mesh, sim.simvars, sim.particles
    driver.Update.update(inp, mesh, sv)
    src = driver.Sourcing.sourcing(mesh, sv, parts, n_census_global=eng.num_particles() * W)
    tr = eng.transport(float(sv.dt), sv.step)
    driver.Clean.clean(parts)
    eng.tally_local()
    if W > 1:
        dep, rad, scalars, kind = imc_dist.reduce_buffer_parts(eng)
        for part in (dep, rad, scalars):
            part.mul_(W)
        torch.cuda.synchronize()
    eng.tally_finish(float(sv.t), float(sv.dt))
    eng.energycheck()
    driver.timestep(str(inp["TIMESTEPPING"]).upper(), sv)
    sv.step += 1
    seg = tr["segments"]                          # read by imc_transport, before the buffer is multiplied
    out.append((tr["variant"], seg, tr["kernel_ms"], seg / tr["kernel_ms"] / 1e6 if tr["kernel_ms"] else 0))
tail = [o for o in out[3:] if o[0] == 2]
print(json.dumps({"world": W, "stripe": os.environ.get("IMC_STRIPE", "default"), "steps": [(v, s, round(k, 2), round(r, 2)) for v, s, k, r in out],
                  "refill_Mseg_per_ms_mean": sum(o[3] for o in tail) / max(len(tail), 1), "kernel_ms_last": out[-1][2], "seg_last": out[-1][1]}))
