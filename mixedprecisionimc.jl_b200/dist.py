"""Particle-sharded multi-GPU time step: one process per GPU, torch.distributed for the plumbing.

The transport step shards by particle: every rank holds the whole mesh, emits the
new-particle ordinals j with j % world == rank, tracks and cleans its own slice, and the per-cell tallies
are combined by an all-reduce of the engine's reduce buffer [energydep | radenergydens | lostenergy | counters] per
time step (issued in two parts so that the large one overlaps compaction and the census tally); the per-cell update
that follows is replicated, so
every rank ends the step with identical temperature / Fleck fields.  The census count needed by the NMAX
cap (imc_sourcing.jl:133-136) is a second, scalar all-reduce.

Backends: NCCL over NVLink on the GPU box (the reduce buffer is wrapped zero-copy as a CUDA tensor);
gloo on CPU for tests with the oracle (the reduce buffer is host memory).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import driver as _driver
from . import lib as _lib


class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def reduce_buffer_parts(engine: _lib.Engine):
    """The engine's reduce buffer as two tensors that share its memory: the deposits [energydep Nc*Ns] and the tail
    [radenergydens Nc | lostenergy | counters].  Element types follow imc_reduce_buffer's `kind`: 0 = Float64 throughout,
    1 = int64 throughout (FIXED tallies), 2 = Float32 deposits (the region's first Nc*Ns 4-byte words: Float16 / Float32
    decks with ATOMIC tallies in global memory accumulate in the deck's own width or wider, like the reference's
    `energydep[cell] += dep`) followed by a Float64 tail at byte offset 8*Nc*Ns."""
    ptr, n, kind = engine.reduce_buffer()
    n_dep = engine.nc * engine.ns
    if engine.lib.backend.startswith("cuda"):
        dev = f"cuda:{engine.cfg.device}"
        t8 = "<i8" if kind == 1 else "<f8"
        dep = torch.as_tensor(_DevArray(ptr, n_dep, "<f4" if kind == 2 else t8), device=dev)
        tail = torch.as_tensor(_DevArray(ptr + 8 * n_dep, n - n_dep, t8), device=dev)
        return dep, tail
    ctype = C.c_int64 if kind == 1 else C.c_double
    arr = torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)))
    if kind == 2:
        dep = torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n_dep,)))
        return dep, arr[n_dep:]
    return arr[:n_dep], arr[n_dep:]


def advance_sharded(sim: _driver.Simulation, group=None) -> dict:
    """One iteration of the reference's while loop (MixedPrecisionIMC.jl:136-152) on a particle shard."""
    inputs, mesh, sv, parts = sim.inputs, sim.mesh, sim.simvars, sim.particles
    eng = mesh.engine
    on_gpu = eng.lib.backend.startswith("cuda")
    rec = {"t": float(sv.t), "dt": float(sv.dt), "step": sv.step}
    _driver.Update.update(inputs, mesh, sv)
    cnt = torch.tensor([eng.num_particles()], dtype=torch.int64, device=f"cuda:{eng.cfg.device}" if on_gpu else "cpu")
    dist.all_reduce(cnt, group=group)
    rec["source"] = _driver.Sourcing.sourcing(mesh, sv, parts, n_census_global=int(cnt.item()))
    rec["transport"] = eng.transport(float(sv.dt), sv.step)
    # The deposits [energydep Nc*Ns] are final once the tracking kernel has returned: their all-reduce starts now, on the
    # collective's stream, and overlaps the compaction and the census tally (which only write the particle list and the
    # [radenergydens | scalars] tail of the buffer); the tail is reduced afterwards.  Two collectives per step, same bytes.
    # (The oracle fills its host-side buffer only in tally_local, so there the whole buffer is reduced afterwards.)
    if on_gpu:
        dep, tail = reduce_buffer_parts(eng)
        work = dist.all_reduce(dep, group=group, async_op=True)
        _driver.Clean.clean(parts)
        eng.tally_local()
        # tally_local only enqueues the census tally on the engine's own stream, which the collective's stream does not
        # know about: imc_reduce_buffer waits for that stream, so the tail is complete before it is reduced
        eng.reduce_buffer()
        dist.all_reduce(tail, group=group)
        work.wait()
        torch.cuda.current_stream().synchronize()
    else:
        _driver.Clean.clean(parts)
        eng.tally_local()
        for part in reduce_buffer_parts(eng):
            dist.all_reduce(part, group=group)
    rec["tally"] = eng.tally_finish(float(sv.t), float(sv.dt))
    rec["energy"] = eng.energycheck()
    _driver.timestep(str(inputs["TIMESTEPPING"]).upper(), sv)
    sv.step += 1
    sim.log.append(rec)
    return rec
