"""How far the shared deterministic elementary functions are from an independent libm, in effect.

The engine, the oracle and the Python restatement all take exp / expm1 / sincos / atan2 / log / pow from csrc/imc_math.h, so
"engine == oracle, bit for bit" cannot see an error in those functions.  The oracle can be built around glibc instead
(IMC_ORACLE_MATH=libm, oracle/imc_oracle_capi.cpp): this file runs the same decks, same seed and Philox draws, once with
each math and bounds the divergence — on the CPU oracle-det vs oracle-libm, and on the GPU the CUDA engine vs oracle-libm.
Stated budget: event flips (a particle whose fate or segment count differs) may occur, because events are decided by exact
floating-point comparisons, but must stay below 0.1 % of the histories; the segment totals within 0.1 %; survivors' values
within a few hundred ulp (ulp-level differences per segment accumulate over a history); tallied fields within 1e-5
(Float32) / 1e-12 (Float64) / 1e-2 (Float16) relative L2, i.e. far inside the Monte Carlo noise (3 sigma ~ 1e-2 here).
(Julia's own Base functions are a third implementation, < 1 ulp like glibc's)"""
import os

import numpy as np
import pytest

from mpimc_b200 import decks, driver, lib

CASES = {
    "suolson-f32": (lambda: decks.suolson(precision="FLOAT32", n_input=20000, n_max=200000), 8, 1e-5, 2.0 ** -23),
    "suolson-f64": (lambda: decks.suolson(precision="FLOAT64", n_input=20000, n_max=200000), 8, 1e-12, 2.0 ** -52),
    "suolson-f16": (lambda: decks.suolson(precision="FLOAT16", n_input=20000, n_max=200000), 8, 1e-2, 2.0 ** -10),
    "crooked-f32": (lambda: decks.crooked_pipe(precision="FLOAT32", n_input=20000, n_max=100000, cellmin=1, pairwise="FALSE"), 6, 1e-5, 2.0 ** -23),
    "crooked-f64": (lambda: decks.crooked_pipe(precision="FLOAT64", n_input=20000, n_max=100000, cellmin=1, pairwise="FALSE"), 6, 1e-12, 2.0 ** -52),
    "marshak-rw-f32": (lambda: decks.marshak(precision="FLOAT32", n_cells=128, nonuniform=True, randomwalk="TRUE", n_input=20000, n_max=100000), 6, 1e-5, 2.0 ** -23),
}


def _run(inputs, library, steps, math=None, **cfg):
    old = os.environ.get("IMC_ORACLE_MATH")
    if math:
        os.environ["IMC_ORACLE_MATH"] = math      # read by the oracle's imc_create
    try:
        sim = driver.setup(inputs, library, **cfg)
    finally:
        if math:
            os.environ.pop("IMC_ORACLE_MATH", None)
            if old is not None:
                os.environ["IMC_ORACLE_MATH"] = old
    sim.save_history = False
    recs = [sim.advance() for _ in range(steps)]
    p, ids = sim.engine.particles()
    return sim, p, ids, recs


def _compare(a, b, field_tol, ulp):
    sa, pa, ia, ra = a
    sb, pb, ib, rb = b
    seg_a, seg_b = (sum(r["transport"]["segments"] for r in x) for x in (ra, rb))
    hist = sum(r["transport"]["histories"] for r in ra)
    assert abs(seg_a - seg_b) <= 1e-3 * seg_a, (seg_a, seg_b)
    for k in ("n_census", "n_absorbed", "n_escaped"):
        ca, cb = (sum(r["transport"][k] for r in x) for x in (ra, rb))
        assert abs(ca - cb) <= 1e-3 * hist + 2, (k, ca, cb)
    common, ka, kb = np.intersect1d(ia, ib, return_indices=True)
    assert len(ia) + len(ib) - 2 * len(common) <= 1e-3 * max(len(ia), 1) + 2      # survivors on one side only
    if len(common):
        x, y = pa[ka], pb[kb]
        same_cell = np.all(x[:, 1:3] == y[:, 1:3], axis=1) if x.shape[1] == 10 else x[:, 2] == y[:, 2]
        assert np.mean(same_cell) >= 0.999
        cols = [7] if x.shape[1] == 10 else [6]                                     # energy slot of the 2-D / 1-D layout
        rel = np.abs(x[same_cell][:, cols] - y[same_cell][:, cols]) / np.maximum(np.abs(y[same_cell][:, cols]), 1e-300)
        assert np.quantile(rel, 0.999) <= 512 * ulp, float(rel.max())
    for name in ("temp", "matenergydens", "radenergydens", "energydep"):
        fa, fb = sa.engine.field(name).astype(np.float64), sb.engine.field(name).astype(np.float64)
        assert np.linalg.norm(fa - fb) <= field_tol * max(np.linalg.norm(fb), 1e-300), name


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_det_math_vs_libm(oracle_lib, case):
    mk, steps, tol, ulp = CASES[case]
    _compare(_run(mk(), oracle_lib, steps), _run(mk(), oracle_lib, steps, math="libm"), tol, ulp)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_cuda_engine_vs_libm_oracle(gpu_lib, oracle_lib, case):
    mk, steps, tol, ulp = CASES[case]
    # EXACT tallies: the engine's sums then follow the reference's order, so the field differences are the math's alone
    _compare(_run(mk(), gpu_lib, steps, tally_mode=lib.TALLY_EXACT), _run(mk(), oracle_lib, steps, math="libm"), tol, ulp)
