#!/usr/bin/env python
"""Benchmark of the IMC transport step (BASELINE.json metric: tracked particle-segments/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path — update -> source -> track -> clean -> tally -> energycheck — over
the workload's particle population.  Workloads (synthetic decks from mixedprecisionimc.jl_b200/decks.py):

  crookedpipe_f32   (default) CrookedPipe 2-D Float32 on the 4096 x 4096 graded mesh, 1.25e8 particles per GPU:
                    BASELINE config 5 in its weak-scaling form, the configuration the north-star target is
                    quoted on (1e9 particles over 8 GPUs).
  crookedpipe_f64   BASELINE config 3 (1024 x 1024, Float64, 1e8 particles).
  marshak_f32_rw    BASELINE config 2 (Marshak 1-D Float32, 2048 graded cells, RANDOMWALK, 1e7 particles).
  suolson_f32       Su-Olson 1-D Float32 scaled to 1e8 particles (config 4 family); suolson_f16 / suolson_f64: its
                    Float16 (ENERGYSCALES 32768, counts kept as integers: Q10) and Float64 members.

`value`  = whole-job segments/s with everything resident in HBM (all stages of the step timed).
`e2e`    = the same through the host-buffer path a stateless drop-in shim uses: every step uploads the
           material state (temp, matenergydens, radenergydens) from pinned host memory and downloads
           the three fields the reference's host reads after the tally, inside the timed region.
`roofline` = tracking kernel only: algorithmic bytes/segment x segments / CUDA-event time of
           the kernel, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline` = the oracle (C++ restatement of the reference, 1 thread like the reference) on a bounded
           sample of the same workload.
--impl reference times that oracle as the reference arm (Julia is not installable here): once on one
           core, the way the reference runs, and once on all host cores it can use (<= 32 processes, particles
           sharded like the engine shards them over GPUs, tallies all-reduced over gloo) — the line's `value` is
           the all-core number, `cpu_baseline.single_core_value` the other.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (deck factory kwargs, particles per GPU, bytes/segment params)
    "crookedpipe_f32": dict(deck="crooked_pipe", precision="FLOAT32", mesh=(4096, 4096), particles=125_000_000, geom=2, s=4),
    "crookedpipe_f64": dict(deck="crooked_pipe", precision="FLOAT64", mesh=(1024, 1024), particles=100_000_000, geom=2, s=8),
    "marshak_f32_rw": dict(deck="marshak", precision="FLOAT32", mesh=(2048,), particles=10_000_000, geom=1, s=4),
    "suolson_f32": dict(deck="suolson", precision="FLOAT32", mesh=(1000,), particles=100_000_000, geom=1, s=4),
    "suolson_f16": dict(deck="suolson", precision="FLOAT16", mesh=(1000,), particles=100_000_000, geom=1, s=2),
    "suolson_f64": dict(deck="suolson", precision="FLOAT64", mesh=(1000,), particles=100_000_000, geom=1, s=8),
}


def workload_cellmin(w: dict, world: int = 1, scaling: str = "weak") -> int:
    """CELLMIN of the synthetic deck.  Crooked pipe: every cell emits at least CELLMIN particles per step (Q9), so under weak
    scaling it grows with the GPU count to keep the per-GPU work fixed (the shipped deck has CELLMIN 10 on 106 x 47 cells;
    on 4096^2 cells that alone would be 1.7e8 particles per step); under strong scaling it stays 1."""
    if w["deck"] == "crooked_pipe":
        return max(1, world) if scaling == "weak" else 1
    return 5 if w["deck"] == "marshak" else 1


def make_inputs(w: dict, particles: int, mesh, world: int = 1, pairwise: str = "FALSE", scaling: str = "weak"):
    """particles = the deck's NMAX (global); NINPUT = NMAX / 2."""
    from mpimc_b200 import decks
    n_max = int(particles)
    n_input = max(n_max // 2, 1)
    cellmin = workload_cellmin(w, world, scaling)
    if w["deck"] == "crooked_pipe":
        es = dict(energyscales=(1024.0,)) if w["precision"] == "FLOAT16" else {}
        return decks.crooked_pipe(precision=w["precision"], n_input=n_input, n_max=n_max, cellmin=cellmin,
                                  mesh_cells=mesh, pairwise=pairwise, **es)
    if w["deck"] == "marshak":
        return decks.marshak(precision=w["precision"], n_cells=mesh[0], nonuniform=True, randomwalk="TRUE", n_input=n_input,
                             n_max=n_max, cellmin=cellmin, pairwise=pairwise)
    if w["deck"] == "suolson":
        return decks.suolson(precision=w["precision"], n_input=n_input, n_max=n_max, pairwise=pairwise)
    raise ValueError(w["deck"])


def cpu_sample_mesh(w: dict, mesh, particles: int, sample: int):
    """Mesh of the bounded CPU sample: 1-D decks keep theirs; the 2-D mesh is scaled so that the sample has the GPU workload's
    particles per cell (cells ~ sample / particles), because every cell emits at least CELLMIN particles per step and the
    number of face crossings per history follows the resolution."""
    if w["geom"] == 1:
        return tuple(mesh)
    f = min(1.0, (sample / max(particles, 1)) ** 0.5)
    return (max(256, int(mesh[0] * f)), max(256, int(mesh[1] * f)))


def bytes_per_segment(geom: int, s: int, seg_per_hist: float) -> float:
    """B = B_seg + B_hist / (segments per history); 64-bit particle id carried (+16)."""
    if geom == 1:
        return 6 * s + (2 * (5 * s + 5) + 16) / max(seg_per_hist, 1e-9)
    return 7 * s + (2 * (6 * s + 9) + 16) / max(seg_per_hist, 1e-9)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 3 + k and s[3 + k] == "Active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _cpu_shard_worker(rank, world, store_path, workload, mesh, sample, steps, warmup, q, pairwise="FALSE"):
    """One shard of the multi-process CPU run: the oracle behind the particle-sharded step of dist.py over gloo."""
    # under torchrun the parent's environment would steer this private group to the elastic agent's store
    for k in [k for k in os.environ if k.startswith("TORCHELASTIC") or k in ("MASTER_ADDR", "MASTER_PORT", "RANK", "WORLD_SIZE", "LOCAL_RANK",
                                                                             "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "GROUP_WORLD_SIZE", "ROLE_WORLD_SIZE")]:
        os.environ.pop(k, None)
    import datetime
    import torch.distributed as dist
    import __graft_entry__ as entry
    from mpimc_b200 import driver, lib
    from mpimc_b200 import dist as imc_dist
    dist.init_process_group("gloo", init_method=f"file://{store_path}", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
    sim = driver.setup(make_inputs(WORKLOADS[workload], sample, mesh, pairwise=pairwise), lib.ImcLib(entry.ORACLE_LIB), rank=rank, world=world)
    sim.save_history = False
    for _ in range(warmup):
        imc_dist.advance_sharded(sim)
    dist.barrier()
    t0 = time.perf_counter()
    seg = 0
    for _ in range(steps):
        seg += imc_dist.advance_sharded(sim)["transport"]["segments"]
    dist.barrier()
    q.put((rank, seg, time.perf_counter() - t0))
    dist.destroy_process_group()


def cpu_port_run_parallel(workload, mesh, sample, steps, warmup, procs, pairwise="FALSE"):
    """The oracle on `procs` host cores: one process per core, particles sharded exactly as the engine shards them over
    GPUs (striped emission, one all-reduce of the tallies per step over gloo).  Returns (segments/s, seconds, segments)."""
    import tempfile
    import torch.multiprocessing as mp
    import __graft_entry__ as entry
    if not os.path.exists(entry.ORACLE_LIB):
        entry.build_oracle()
    tmp = tempfile.mkdtemp(prefix="imc_cpu_")
    store_path = os.path.join(tmp, "store")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_cpu_shard_worker, args=(r, procs, store_path, workload, mesh, sample, steps, warmup, q, pairwise)) for r in range(procs)]
    for p_ in ps:
        p_.start()
    try:
        res = [q.get(timeout=240) for _ in range(procs)]   # a stuck rendezvous must not stall the bench: the caller falls back
    finally:
        for p_ in ps:
            p_.join(timeout=10)
            if p_.is_alive():
                p_.kill()
        try:
            if os.path.exists(store_path):
                os.remove(store_path)
            os.rmdir(tmp)
        except OSError:
            pass
    seg = sum(r[1] for r in res)
    dt = max(r[2] for r in res)
    return seg / dt, dt, seg


def traffic_from_profile(workload, mesh, particles, segments_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the tracking kernel per launch.  ncu cannot run inside the timed bench,
    so the figure is the committed `ncu --set full` capture of this workload (profiles/traffic.json: DRAM bytes and segments
    of the captured launch) scaled to this run's segments per launch; None when no capture of this workload / mesh exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        for e in json.load(open(p)):
            if e["workload"] == workload and list(e["mesh"]) == list(mesh) and e["particles_per_gpu"] == particles:
                return e["dram_bytes_per_launch"] / e["segments_per_launch"] * segments_per_launch
    except Exception:
        pass
    return None


def tally_rel_err(w, mesh, sample_particles: int, glib, tally_mode, device: int):
    """BASELINE.json's second metric, "tally rel. err vs reference": one time step of the same bounded sample on the
    engine (the bench's tally mode) and on the oracle, from identical state and with identical draws — the particles
    are then bit-identical, so the difference of the tallied fields is the engine's summation order (atomics) and its
    reciprocal-form deposits.  Relative L2 error per field, and the Float64 energy-balance residual of each side."""
    import __graft_entry__ as entry
    from mpimc_b200 import driver, lib
    olib = lib.ImcLib(entry.ORACLE_LIB)
    inputs = make_inputs(w, sample_particles, mesh)
    a = driver.setup(inputs, glib, device=device, tally_mode=tally_mode)
    b = driver.setup(inputs, olib)
    a.save_history = b.save_history = False
    ra, rb = a.advance(), b.advance()
    out = {"sample": f"1 step, {sample_particles} particles (NMAX), mesh {'x'.join(map(str, mesh))}, same seed on both sides",
           "segments_equal": ra["transport"]["segments"] == rb["transport"]["segments"]}
    for name in ("energydep", "radenergydens", "matenergydens", "temp"):
        fa, fb = a.engine.field(name).astype(np.float64), b.engine.field(name).astype(np.float64)
        out[name] = float(np.linalg.norm(fa - fb) / max(np.linalg.norm(fb), 1e-300))
    out["energy_error_engine"], out["energy_error_reference_port"] = ra["energy"]["energy_error"], rb["energy"]["energy_error"]
    return out


def cpu_port_run(w, mesh, sample_particles: int, steps: int, warmup: int, pairwise: str = "FALSE"):
    """Time the oracle (1 thread, like the reference) on a bounded sample of the workload."""
    import __graft_entry__ as entry
    from mpimc_b200 import driver, lib
    if not os.path.exists(entry.ORACLE_LIB):
        entry.build_oracle()
    olib = lib.ImcLib(entry.ORACLE_LIB)
    inputs = make_inputs(w, sample_particles, mesh, pairwise=pairwise)
    sim = driver.setup(inputs, olib)
    sim.save_history = False
    for _ in range(warmup):
        sim.advance()
    seg = 0
    hist = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        r = sim.advance()
        seg += r["transport"]["segments"]
        hist += r["transport"]["histories"]
    dt = time.perf_counter() - t0
    return seg / dt, dt, seg, hist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="crookedpipe_f32", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (default: the workload's)")
    ap.add_argument("--mesh", type=int, nargs="*", default=None, help="cells per axis (default: the workload's)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="particles in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tally", default="auto", choices=["auto", "atomic", "fixed"])
    ap.add_argument("--track", default="auto", choices=["auto", "history", "refill", "event"], help="tracking schedule (auto = measured)")
    ap.add_argument("--pairwise", default="FALSE", choices=["FALSE", "TRUE"], help="the deck's PAIRWISE keyword (CrookedPipe.txt:96 ships TRUE)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --particles per GPU (NMAX grows with the GPU count); strong: --global-particles in total, split over the GPUs")
    ap.add_argument("--global-particles", type=int, default=1_000_000_000, help="NMAX of the strong-scaling run (BASELINE config 5: 1e9)")
    args = ap.parse_args()

    w = WORKLOADS[args.workload]
    mesh = tuple(args.mesh) if args.mesh else w["mesh"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.scaling == "strong":
        n_global = args.global_particles
        particles = n_global // world
    else:
        particles = args.particles or w["particles"]
        n_global = particles * world
    cellmin = workload_cellmin(w, world, args.scaling)
    config = {"workload": f"{args.workload}: {w['deck']} {w['precision']} mesh {'x'.join(map(str, mesh))}, NMAX = {n_global} particles "
                          f"({particles} per GPU; NMAX caps the census + source population, the histories tracked per step are reported "
                          f"as histories_per_step), NINPUT = NMAX/2, CELLMIN {cellmin}, PAIRWISE {args.pairwise}",
              "mesh": list(mesh), "particles_per_gpu": particles, "nmax_global": n_global, "cellmin": cellmin, "pairwise": args.pairwise,
              "precision": w["precision"], "scaling_mode": args.scaling,
              "l2_policy": "inputs larger than L2 (particle state >> 126 MB); no explicit flush",
              "tally_mode": args.tally, "tracking": f"schedule {args.track} (history static grid-stride | history warp-refill | event-based), 256 threads/block"}

    # ---------------------------------------------------------------- reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return
        sample = args.cpu_sample or 10_000_000
        cmesh = cpu_sample_mesh(w, mesh, particles, sample)
        nsteps, nwarm = max(1, min(args.steps, 3)), min(args.warmup, 1)
        # The reference itself is single-threaded Julia (no Threads / Distributed anywhere in the package); its C++ port is
        # timed here on all host cores it can use by sharding the particles over processes the way the engine shards them
        # over GPUs, and on one core the way the reference runs.
        cores = max(1, min(os.cpu_count() or 1, 32))
        seg_1, dt_1, seg1, _ = cpu_port_run(w, cmesh, sample, nsteps, nwarm, args.pairwise)
        seg_s, dt, seg, used, psample, pmesh = seg_1, dt_1, seg1, 1, sample, cmesh
        if cores > 1:
            try:
                psample = sample * max(1, cores // 8)   # keep >= 1 s of work per timed step on a many-core host
                pmesh = cpu_sample_mesh(w, mesh, particles, psample)
                seg_s, dt, seg = cpu_port_run_parallel(args.workload, pmesh, psample, nsteps, nwarm, cores, args.pairwise)
                used = cores
            except Exception as e:  # keep the single-core number rather than lose the arm
                sys.stderr.write(f"bench.py: multi-process CPU run failed ({e}); reporting the single-core run\n")
                psample, pmesh = sample, cmesh
        # this arm's OWN configuration: a bounded sample of the workload (same deck, same physics, mesh scaled to keep the
        # workload's particles per cell), not the GPU arm's mesh and population
        rconfig = dict(config)
        rconfig.update({"workload": f"{args.workload}: bounded CPU sample of that workload — {w['deck']} {w['precision']} mesh "
                                    f"{'x'.join(map(str, pmesh))}, NMAX = {psample} particles, NINPUT = NMAX/2, CELLMIN {workload_cellmin(w)}, "
                                    f"PAIRWISE {args.pairwise}, {nsteps} steps after {nwarm} warm-up from t = 0",
                        "mesh": list(pmesh), "particles_per_gpu": None, "nmax_global": psample, "cellmin": workload_cellmin(w),
                        "gpu_arm_mesh": list(mesh), "gpu_arm_nmax_global": n_global,
                        "tally_mode": "reference order (sequential += / Base.sum)", "tracking": "history-based, one thread per process"})
        line = {"metric": "tracked particle-segments/sec", "value": seg_s, "unit": "segments/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / nsteps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": w["precision"].lower().replace("float", "f"),
                "data": "synthetic", "config": rconfig, "impl": "reference",
                "cpu_baseline": {"value": seg_s, "unit": "segments/s", "cores": used, "kind": "port", "single_core_value": seg_1,
                                 "sample": f"oracle (C++ restatement of the single-threaded Julia reference) on {psample} particles, mesh "
                                           f"{'x'.join(map(str, pmesh))}, "
                                           f"{nsteps} steps after {nwarm} warm-up: {seg} segments in {dt:.2f} s on {used} process(es), "
                                           f"particles sharded over processes like the engine shards them over GPUs (gloo all-reduce of the "
                                           f"tallies); one process (the way the single-threaded reference runs) on {sample} particles, mesh "
                                           f"{'x'.join(map(str, cmesh))}: {seg1} segments in {dt_1:.2f} s"},
                "e2e": {"value": seg_s, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    from mpimc_b200 import driver, lib
    from mpimc_b200 import dist as imc_dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the transport step has no CPU fallback (use --impl reference for the CPU oracle)")
    if not os.path.exists(entry.LIB):
        entry.build_cuda()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    glib = lib.ImcLib(entry.LIB)
    tally_mode = {"auto": lib.TALLY_AUTO, "atomic": lib.TALLY_ATOMIC, "fixed": lib.TALLY_FIXED}[args.tally]
    inputs = make_inputs(w, n_global, mesh, world, args.pairwise, args.scaling)  # NMAX / NINPUT are global; each rank emits its stripe
    track_mode = {"auto": lib.TRACK_AUTO, "history": lib.TRACK_HISTORY, "refill": lib.TRACK_REFILL, "event": lib.TRACK_EVENT}[args.track]
    sim = driver.setup(inputs, glib, device=local_rank, rank=rank, world=world, tally_mode=tally_mode, track_mode=track_mode)
    sim.save_history = False
    eng = sim.engine
    nc = eng.nc
    dev = torch.device(f"cuda:{local_rank}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if world > 1:
            return imc_dist.advance_sharded(sim)
        return sim.advance()

    # pinned host buffers for the end-to-end path, in the fields' own element type (Array{T}, as the Julia shim
    # passes pointer(mesh.x)); mesh.temp is 8-byte once a LINEARIZED deck has made it Float64 (Q12)
    sim_dt = {k: eng.field_dtype(k) for k in ("temp", "matenergydens", "radenergydens")}
    tdt = {np.float16: torch.float16, np.float32: torch.float32, np.float64: torch.float64}
    pin = {k: torch.empty(nc, dtype=tdt[sim_dt[k]]).pin_memory().numpy() for k in sim_dt}

    dev_state = {}

    def step_e2e():
        if world == 1:
            eng.set_state_native(temp=pin["temp"], matenergydens=pin["matenergydens"], radenergydens=pin["radenergydens"])  # H2D
            r = step_resident()
            for k in pin:                                                                                                   # D2H
                eng.field_native(k, out=pin[k])
            return r
        # N GPUs: the per-cell state is replicated, so it crosses PCIe ONCE (rank 0, from pinned memory), is broadcast over
        # NVLink, and every engine takes it from device memory; rank 0 alone reads the step's result back
        for k in pin:
            t = dev_state.get(k)
            if t is None or t.dtype != tdt[pin[k].dtype.type]:
                t = dev_state[k] = torch.empty(nc, dtype=tdt[pin[k].dtype.type], device=dev)
            if rank == 0:
                t.copy_(torch.from_numpy(pin[k]), non_blocking=True)                                                        # H2D
            dist.broadcast(t, src=0)
        torch.cuda.current_stream().synchronize()
        eng.set_state_native(temp=dev_state["temp"], matenergydens=dev_state["matenergydens"], radenergydens=dev_state["radenergydens"])
        r = step_resident()
        if rank == 0:
            for k in pin:                                                                                                   # D2H
                eng.field_native(k, out=pin[k])
        return r

    import copy
    for _ in range(args.warmup):
        step_resident()
    barrier()
    # Restart point: the resident loop and the host-buffer (e2e) loop below both run time steps W .. W+K-1 of the same
    # simulation from this state (imc_checkpoint copies the engine's state on the device; the host's t / dt / step are
    # copied here), so their step times compare one to one.
    if pin["temp"].dtype != eng.field_dtype("temp"):   # mesh.temp turned Float64 during the warm-up (Q12: first LINEARIZED tally)
        pin["temp"] = torch.empty(nc, dtype=tdt[eng.field_dtype("temp")]).pin_memory().numpy()
    for k in pin:
        eng.field_native(k, out=pin[k])
    eng.checkpoint("save")
    sv0 = copy.deepcopy(sim.simvars)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.kernel_launches()
    seg = hist = 0
    kms = 0.0
    variants = []
    modes = set()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    estream = torch.cuda.ExternalStream(eng.stream(), device=dev)   # the stream the engine launches its kernels on
    barrier()
    ev0.record(estream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = step_resident()
        seg += r["transport"]["segments"]; hist += r["transport"]["histories"]; kms += r["transport"]["kernel_ms"]
        variants.append({1: "static", 2: "refill", 3: "event"}.get(r["transport"]["variant"], "?"))
        modes.add({0: "auto", 1: "atomic", 2: "fixed", 3: "exact"}.get(r["transport"].get("tally_mode", -1), "?"))
    ev1.record(estream)
    barrier()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) * 1e-3          # device time of the K steps on the engine's stream
    launches = eng.kernel_launches() - l0
    n_part = eng.num_particles()
    # end-to-end: the SAME K time steps again, through host buffers
    eng.checkpoint("restore")
    sim.simvars = copy.deepcopy(sv0)
    if pin["temp"].dtype != eng.field_dtype("temp"):   # element type of mesh.temp at the restart point (Q12)
        pin["temp"] = torch.empty(nc, dtype=tdt[eng.field_dtype("temp")]).pin_memory().numpy()
        eng.field_native("temp", out=pin["temp"])
    barrier()
    ev2.record(estream)
    t1 = time.perf_counter()
    seg_e = 0
    for _ in range(args.steps):
        if pin["temp"].dtype != eng.field_dtype("temp"):   # mesh.temp turned Float64 in the previous step (Q12, first LINEARIZED tally)
            pin["temp"] = torch.empty(nc, dtype=torch.float64).pin_memory().numpy()
            eng.field_native("temp", out=pin["temp"])
        r = step_e2e()
        seg_e += r["transport"]["segments"]
    ev3.record(estream)
    barrier()
    wall_e = time.perf_counter() - t1
    dev_e_s = ev2.elapsed_time(ev3) * 1e-3
    eng.checkpoint("drop")
    io_bytes = sum(int(v.nbytes) for v in pin.values())
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    tot = torch.tensor([seg, hist, seg_e, n_part], dtype=torch.float64, device=dev)
    mx = torch.tensor([dev_s, dev_e_s, kms, wall, wall_e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    kms_all = [kms]
    if world > 1:   # per-rank tracking-kernel time: the max enters the step time, the spread says how much is rank imbalance
        gathered = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, torch.tensor([kms], dtype=torch.float64, device=dev))
        kms_all = [float(g.item()) for g in gathered]
    seg_g, hist_g, seg_e_g, n_part_g = tot.tolist()
    wall_g, wall_e_g, kms_g, host_wall_g, host_wall_e_g = mx.tolist()   # the first two are CUDA-event times (max over ranks)

    if rank == 0:
        peak, peak_src = hbm_peak()
        sph = seg_g / max(hist_g, 1)
        bps = bytes_per_segment(w["geom"], w["s"], sph)
        # dominant kernel: tracking.  per launch: this rank's segments x bytes/segment over its event time
        ach = (seg * bps) / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
        line = {
            "metric": "tracked particle-segments/sec", "value": seg_g / wall_g, "unit": "segments/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall_g / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": w["precision"].lower().replace("float", "f"), "data": "synthetic",
            "config": config,
            "histories_per_s": hist_g / wall_g, "histories_per_step": hist_g / args.steps, "segments_per_step": seg_g / args.steps,
            "segments_per_history": sph, "particles_resident": n_part_g, "tally_modes_run": sorted(modes),
            "tracking_kernel_ms_per_step": kms_g / args.steps, "tracking_kernel_share_of_step": kms_g / (1e3 * wall_g),
            "tracking_kernel_ms_per_step_by_rank": [k / args.steps for k in kms_all],
            "timing": "CUDA events on the engine's stream, max over ranks", "host_wall_ms_per_step": 1e3 * host_wall_g / args.steps,
            "e2e": {"value": seg_e_g / wall_e_g, "unit": "segments/s", "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes,
                    "ms_per_step": 1e3 * wall_e_g / args.steps, "host_wall_ms_per_step": 1e3 * host_wall_e_g / args.steps,
                    "segments_per_step": seg_e_g / args.steps,
                    "path": "imc_set_state_native (pinned Array{T} -> device) + the step + imc_get_field_native x3 (device -> pinned Array{T}); "
                            "the same K time steps as `value`, restarted from imc_checkpoint"
                            + ("; N > 1: the replicated state crosses PCIe once on rank 0 and is broadcast over NVLink, rank 0 reads the result" if world > 1 else "")},
            "gpu_launches": launches, "schedule_per_step": variants,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic_from_profile(args.workload, mesh, particles, seg / max(args.steps, 1)),
                         "traffic_source": "profiles/traffic.json: dram__bytes of the committed ncu capture of this kernel, scaled by segments per launch",
                         "kernel": ("k_track1d_rw" if w["deck"] == "marshak" else
                                    {"refill": f"k_track_refill<{w['geom']}-D>", "static": f"k_track{w['geom']}d", "event": "k_track_event"}.get(variants[-1], "?")),
                         "bytes_per_segment": bps, "peak_source": peak_src},
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline and world == 1:   # the CPU legs run on rank 0 at N = 1 only
            sample = args.cpu_sample or 4_000_000
            cmesh = cpu_sample_mesh(w, mesh, particles, sample)
            seg_s, dt, cseg, chist = cpu_port_run(w, cmesh, sample, 2, 1, args.pairwise)
            line["cpu_baseline"] = {"value": seg_s, "unit": "segments/s", "cores": 1, "kind": "port",
                                    "sample": f"oracle (C++ restatement of the Julia reference, single-threaded like it) on {sample} particles, "
                                              f"mesh {'x'.join(map(str, cmesh))}, 2 steps after 1 warm-up: {cseg} segments in {dt:.1f} s"}
            try:
                line["tally_rel_err"] = tally_rel_err(w, cpu_sample_mesh(w, mesh, particles, min(sample, 2_000_000)), min(sample, 2_000_000), glib, tally_mode, local_rank)
            except Exception as e:  # a reported metric, never a reason to lose the bench line
                line["tally_rel_err"] = {"error": str(e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
