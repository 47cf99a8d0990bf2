"""Particle-sharded time step on TWO GPUs over NCCL (dist.py) against the one-GPU run of the same deck.

With the FIXED tally mode (64-bit fixed-point accumulators, summed as integers by the all-reduce) the result
must not depend on the number of GPUs: the union of the two ranks' particles, matched by particle id, and every
field are BIT-IDENTICAL to the single-GPU run, step after step.  Skipped on a one-GPU box (the host logic is
covered on CPU by tests/test_dist_gloo.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _deck(decks, name):
    if name == "suolson":
        return decks.suolson(precision="FLOAT32", n_input=30000, n_max=120000, pairwise="FALSE")   # NMAX cap becomes active
    return decks.crooked_pipe(precision="FLOAT32", n_input=30000, n_max=600000, cellmin=1, pairwise="FALSE")


def _worker(rank, world, port, deckname, steps, q, tally="fixed"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from mpimc_b200 import decks, driver, lib
    from mpimc_b200 import dist as imc_dist
    glib = lib.ImcLib(entry.LIB)
    sim = driver.setup(_deck(decks, deckname), glib, device=rank, rank=rank, world=world,
                       tally_mode={"fixed": lib.TALLY_FIXED, "atomic": lib.TALLY_ATOMIC, "auto": lib.TALLY_AUTO}[tally])
    sim.save_history = False
    recs = []
    for _ in range(steps):
        r = imc_dist.advance_sharded(sim)
        recs.append((r["source"]["n_new_global"], r["source"]["n_new_local"], r["source"]["totalenergy"], r["transport"]["segments"]))
    slots, ids = sim.engine.particles()
    fields = {k: sim.engine.field(k) for k in ("temp", "energydep", "radenergydens", "matenergydens", "fleck")}
    q.put((rank, recs, slots, ids, fields))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("deckname", ["suolson", "crooked_pipe"])
def test_two_gpu_run_is_bit_identical_to_one_gpu(built, deckname):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from mpimc_b200 import decks, driver, lib
    steps, world = 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, deckname, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    glib = lib.ImcLib(entry.LIB)
    sim = driver.setup(_deck(decks, deckname), glib, tally_mode=lib.TALLY_FIXED)
    sim.save_history = False
    single = [sim.advance() for _ in range(steps)]
    slots1, ids1 = sim.engine.particles()
    for s in range(steps):
        g = [results[r][1][s] for r in range(world)]
        assert g[0][0] == g[1][0] == single[s]["source"]["n_new_global"]
        assert g[0][1] + g[1][1] == g[0][0] and abs(g[0][1] - g[1][1]) <= 1
        assert g[0][2] == g[1][2] == single[s]["source"]["totalenergy"]
        assert g[0][3] + g[1][3] == single[s]["transport"]["segments"]
    slots = np.concatenate([results[r][2] for r in range(world)])
    ids = np.concatenate([results[r][3] for r in range(world)])
    assert len(np.unique(ids)) == len(ids)
    o, o1 = np.argsort(ids), np.argsort(ids1)
    assert np.array_equal(ids[o], ids1[o1])
    assert np.array_equal(slots[o], slots1[o1])
    for k in ("temp", "energydep", "radenergydens", "matenergydens", "fleck"):
        assert np.array_equal(results[0][4][k], results[1][4][k]), k
        assert np.array_equal(results[0][4][k], sim.engine.field(k)), k


@pytest.mark.parametrize("deckname,tally", [("suolson", "atomic"), ("crooked_pipe", "atomic"), ("crooked_pipe", "auto")])
def test_two_gpu_float_tallies_agree_with_one_gpu(built, deckname, tally):
    """Float (ATOMIC) tallies on two GPUs: the ranks exchange Float32 images of their Float64 per-cell partial sums
    (dist.py::_all_reduce_field), so the fields are not bit-identical to the one-GPU run but must agree with it to Float32
    rounding of sums of a few partial results; both ranks must hold identical fields (the update is replicated), and the
    source statistics — which depend on the last bit of mesh.totalenergy and hence on temp — must stay equal step after step
    as long as the fields agree (first step exactly)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from mpimc_b200 import decks, driver, lib
    steps, world = 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, deckname, steps, q, tally)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    glib = lib.ImcLib(entry.LIB)
    sim = driver.setup(_deck(decks, deckname), glib, tally_mode={"atomic": lib.TALLY_ATOMIC, "auto": lib.TALLY_AUTO}[tally])
    sim.save_history = False
    single = [sim.advance() for _ in range(steps)]
    g = [results[r][1][0] for r in range(world)]
    assert g[0][0] == g[1][0] == single[0]["source"]["n_new_global"] and g[0][3] + g[1][3] == single[0]["transport"]["segments"]
    for k in ("temp", "energydep", "radenergydens", "matenergydens", "fleck"):
        a, b, c = results[0][4][k], results[1][4][k], sim.engine.field(k)
        assert np.array_equal(a, b), k                                               # replicated update: identical on both ranks
        assert np.linalg.norm(a - c) <= 2e-5 * max(np.linalg.norm(c), 1e-300), (k, float(np.linalg.norm(a - c) / np.linalg.norm(c)))
