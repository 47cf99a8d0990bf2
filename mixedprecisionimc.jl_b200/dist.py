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


# Float16 / Float32 decks, float (ATOMIC / EXACT) tallies: exchange the per-cell sums as Float32 (see _all_reduce_field)
NARROW_FIELD_REDUCE = True


class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def reduce_buffer_parts(engine: _lib.Engine):
    """The engine's reduce buffer as (deposits [energydep Nc*Ns], fields [radenergydens Nc], scalars [lostenergy | counters],
    kind): three tensors that share its memory; kind 0 = Float64, 1 = int64 (FIXED tallies) — imc_reduce_buffer."""
    ptr, n, kind = engine.reduce_buffer()
    n_dep, nc = engine.nc * engine.ns, engine.nc
    if engine.lib.backend.startswith("cuda"):
        buf = torch.as_tensor(_DevArray(ptr, n, "<i8" if kind == 1 else "<f8"), device=f"cuda:{engine.cfg.device}")
    else:
        ctype = C.c_int64 if kind == 1 else C.c_double
        buf = torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)))
    return buf[:n_dep], buf[n_dep:n_dep + nc], buf[n_dep + nc:], kind


def _all_reduce_field(part: torch.Tensor, narrow: bool, group=None, async_op: bool = False):
    """Sum one per-cell part of the reduce buffer over the ranks.  narrow: Float16 / Float32 decks with float tallies — the
    sums end up in an Array{T} of that precision anyway (and the reference accumulates them in T, imc_transport.jl:120),
    so the ranks exchange Float32 images of their Float64 partial sums: half the bytes on the wire.  Returns a callable
    that completes the reduction (waits, converts back)."""
    if not narrow:
        work = dist.all_reduce(part, group=group, async_op=async_op)
        return (lambda: work.wait()) if async_op else (lambda: None)
    img = part.to(torch.float32)
    work = dist.all_reduce(img, group=group, async_op=async_op)

    def finish():
        if async_op:
            work.wait()
        part.copy_(img)
    return finish


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
        dep, rad, scalars, kind = reduce_buffer_parts(eng)
        narrow = kind == 0 and eng.cfg.precision != _lib.F64 and NARROW_FIELD_REDUCE
        done_dep = _all_reduce_field(dep, narrow, group, async_op=True)
        _driver.Clean.clean(parts)
        eng.tally_local()
        # tally_local only enqueues the census tally on the engine's own stream, which the collective's stream does not
        # know about: imc_reduce_buffer waits for that stream, so the tail is complete before it is reduced
        eng.reduce_buffer()
        done_rad = _all_reduce_field(rad, narrow, group)
        dist.all_reduce(scalars, group=group)     # lostenergy and the event counters stay 8-byte
        done_rad(); done_dep()
        torch.cuda.current_stream().synchronize()
    else:
        _driver.Clean.clean(parts)
        eng.tally_local()
        dep, rad, scalars, _ = reduce_buffer_parts(eng)
        for part in (dep, rad, scalars):
            dist.all_reduce(part, group=group)
    rec["tally"] = eng.tally_finish(float(sv.t), float(sv.dt))
    rec["energy"] = eng.energycheck()
    _driver.timestep(str(inputs["TIMESTEPPING"]).upper(), sv)
    sv.step += 1
    sim.log.append(rec)
    return rec
