"""Programmatic input decks: the reference's shipped problems (src/inputs/*.txt) restated as dicts in the
format ``deck.read_inputs`` returns, plus the scaled-up variants BASELINE.json's configs ask for
(SURVEY.md §8d).  Nothing here reads /root/reference, so tests and benchmarks run on the GPU box.

Scaled meshes keep every material interface on a mesh node (``region_joiner`` snaps regions with
``findlast(x -> x <= edge, nodes)``, imc_mesh.jl:211-214).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .deck import _PRECISIONS, parse_T


def _T(precision):
    return _PRECISIONS[precision] if isinstance(precision, str) else precision


def _arr(T, values) -> np.ndarray:
    return np.array([parse_T(T, repr(float(v))) for v in values], dtype=T)


def _common(name, precision, seed, geometry):
    return {"NAME": name, "PRECISION": _T(precision), "SEED": str(seed), "GEOMETRY": geometry,
            "PLOTVARS": "RADENERGY", "BENCHMARK": "FALSE", "SAVEFIG": "FALSE", "SAVEANIMATION": "FALSE", "SAVEVARS": "FALSE"}


SUOLSON_XBENCH = [0.01, 0.1, 0.17783, 0.31623, 0.45, 0.5, 0.56234, 0.75, 1.0, 1.33352, 1.77828, 3.16228]
SUOLSON_YBENCH = [0.72799, 0.71888, 0.69974, 0.63203, 0.50315, 0.40769, 0.29612, 0.13756, 0.04396, 0.00324, 0.0, 0.0]


def suolson(precision="FLOAT64", n_input=1000, n_max=50000, pairwise="FALSE", energyscales: Optional[Sequence[float]] = None,
            dx="0.01", xsize="10.0", dt="0.002", endtime="10.0", seed=1234, cellmin=1) -> Dict:
    """Su-Olson linearized benchmark (src/inputs/SuOlson.txt; shipped as FLOAT16 with ENERGYSCALES 32768)."""
    T = _T(precision)
    d = _common("Su Olson Benchmark", precision, seed, "1D")
    if energyscales is None:
        energyscales = [32768.0] if T is np.float16 else [1.0]
    d.update({
        "MESHTYPE": "UNIFORM", "XSIZE": xsize, "DX": dx,
        "SIGMA_A_REGS": [xsize], "SIGMA_A_VALS": ["0.5"], "SIGMA_A_POWERS": ["0.0"],
        "SIGMA_S_REGS": [xsize], "SIGMA_S_VALS": ["0.5"], "SIGMA_S_POWERS": ["0.0"],
        "RADSOURCE_REGS": _arr(T, [0.5, float(xsize)]), "RADSOURCE_VALS": _arr(T, [1.0, 0.0]),
        "BEE_REGS": [xsize], "BEE_VALS": ["1.0"],
        "LEFTBC": "REFLECT", "RIGHTBC": "VACUUM",
        "TIMESTEPPING": "CONSTANT", "DT": dt, "ENDTIME": endtime,
        "NINPUT": str(n_input), "NMAX": str(n_max), "CELLMIN": str(cellmin),
        "T_INIT": "0.00316", "T_SURFACE_VALS": _arr(T, [0.0, 0.0]), "T_SURFACE_REGS": [""],
        "PHYS_C": "1.0", "PHYS_A": "1.0", "ALPHA": "4.0",
        "LINEARIZED": "TRUE", "PAIRWISE": pairwise, "RANDOMWALK": "FALSE",
        "ENERGYSCALES": _scales(T, energyscales), "DISTANCESCALE": "1",
    })
    return d


def _scales(T, energyscales):
    if len(energyscales) == 1:
        return [repr(float(energyscales[0]))]  # single-element arrays stay strings (Q22)
    return _arr(T, energyscales)


def infinite_medium(precision="FLOAT32", n_input=10000, n_max=50000, pairwise="TRUE", seed=12345, randomwalk="FALSE",
                    energyscales=(1.0,)) -> Dict:
    """src/inputs/InfiniteMedium.txt: equilibrium cv*T + a*T^4 = cv*T0  ->  T_eq = 0.98698."""
    T = _T(precision)
    d = _common("Infinite Medium", precision, seed, "1D")
    d.update({
        "MESHTYPE": "UNIFORM", "XSIZE": "1.0", "DX": "0.05",
        "SIGMA_A_REGS": [""], "SIGMA_A_VALS": ["1000.0"], "SIGMA_A_POWERS": ["0.0"],
        "SIGMA_S_REGS": [""], "SIGMA_S_VALS": ["0.0"], "SIGMA_S_POWERS": ["0.0"],
        "BEE_REGS": [""], "BEE_VALS": ["1.0"], "RADSOURCE_REGS": [""], "RADSOURCE_VALS": ["0"],
        "LEFTBC": "REFLECT", "RIGHTBC": "REFLECT",
        "TIMESTEPPING": "CONSTANT", "DT": "0.0005", "ENDTIME": "1.0",
        "NINPUT": str(n_input), "NMAX": str(n_max), "CELLMIN": "1",
        "T_INIT": "1.00", "T_SURFACE_VALS": _arr(T, [0.0, 0.0]), "T_SURFACE_REGS": [""],
        "PHYS_C": "299.70", "PHYS_A": "0.01372016", "ALPHA": "1.0",
        "LINEARIZED": "FALSE", "PAIRWISE": pairwise, "RANDOMWALK": randomwalk,
        "ENERGYSCALES": _scales(T, list(energyscales)), "DISTANCESCALE": "1",
    })
    return d


def graded_nodes_1d(length: float, n_cells: int, dx_min: float) -> np.ndarray:
    """Nodes on [0, length] with geometric grading toward x = 0 (first cell dx_min), Float64."""
    if n_cells * dx_min >= length:
        return np.linspace(0.0, length, n_cells + 1)
    lo, hi = 1.0 + 1e-12, float(np.exp(600.0 / n_cells))  # keep r ** n_cells finite
    f = lambda r: dx_min * (r ** n_cells - 1.0) / (r - 1.0) - length
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if f(mid) > 0:
            hi = mid
        else:
            lo = mid
    r = 0.5 * (lo + hi)
    widths = dx_min * r ** np.arange(n_cells)
    nodes = np.concatenate([[0.0], np.cumsum(widths)])
    nodes *= length / nodes[-1]
    nodes[-1] = length
    return nodes


def marshak(precision="FLOAT64", n_cells=300, nonuniform=False, randomwalk="FALSE", n_input=10000, n_max=60000,
            cellmin=5, pairwise="TRUE", seed=12345, dx_min=1e-5, energyscales=(1.0,)) -> Dict:
    """Marshak wave (src/inputs/MarshakWave.txt: UNIFORM, 300 cells, FLOAT64, RANDOMWALK FALSE).  BASELINE config 2
    is the derived deck: FLOAT32, 2048 graded cells, RANDOMWALK TRUE, NMAX 1e7."""
    T = _T(precision)
    d = _common("Marshak Wave", precision, seed, "1D")
    if nonuniform:
        d.update({"MESHTYPE": "NONUNIFORM", "MESHNODES": graded_nodes_1d(0.15, n_cells, dx_min)})
    else:
        d.update({"MESHTYPE": "UNIFORM", "XSIZE": "0.15", "DX": repr(0.15 / n_cells)})
    d.update({
        "SIGMA_A_REGS": ["0.15"], "SIGMA_A_VALS": ["1000.0"], "SIGMA_A_POWERS": ["-3.0"],
        "SIGMA_S_REGS": ["0.15"], "SIGMA_S_VALS": ["0.0"], "SIGMA_S_POWERS": ["0.0"],
        "RADSOURCE_REGS": ["0.15"], "RADSOURCE_VALS": ["0.0"], "BEE_REGS": ["0.15"], "BEE_VALS": ["0.3"],
        "LEFTBC": "VACUUM", "RIGHTBC": "VACUUM",
        "TIMESTEPPING": "RAMP", "DT0": "0.00001", "K": "1.01", "DTMAX": "0.001", "ENDTIME": "2.0",
        "NINPUT": str(n_input), "NMAX": str(n_max), "CELLMIN": str(cellmin),
        "T_INIT": "0.01", "T_SURFACE_VALS": _arr(T, [1.0, 0.01]), "T_SURFACE_REGS": [""],
        "PHYS_C": "299.70", "PHYS_A": "0.01372016", "ALPHA": "1.0",
        "LINEARIZED": "FALSE", "PAIRWISE": pairwise, "RANDOMWALK": randomwalk,
        "ENERGYSCALES": _scales(T, list(energyscales)), "DISTANCESCALE": "1.0",
    })
    return d


# ---- Crooked pipe ---------------------------------------------------------------------------------
# Shipped mesh (src/inputs/CrookedPipe.txt:16-47): 0.1-wide cells, graded over the last / first 0.1 next to
# each material interface with ten cells whose widths grow geometrically from 1e-3.
CP_X_EDGES = [2.5, 3.0, 4.0, 4.5]     # interfaces refined on both sides where the deck does
CP_X_REFINE = {2.5: "left", 3.0: "right", 4.0: "left", 4.5: "right"}
CP_Y_REFINE = {0.5: "right", 1.0: "left", 1.5: "right"}


def _graded_unit(n: int = 10, first: float = 1e-3, total: float = 0.1) -> np.ndarray:
    """Offsets 0 < d_1 < ... < d_n = total with geometric widths starting at `first` (the deck's 5-decimal values)."""
    lo, hi = 1.0 + 1e-9, 3.0
    for _ in range(200):
        r = 0.5 * (lo + hi)
        if first * (r ** n - 1) / (r - 1) > total:
            hi = r
        else:
            lo = r
    r = 0.5 * (lo + hi)
    d = first * (r ** np.arange(1, n + 1) - 1) / (r - 1)
    d[-1] = total
    return d


def crooked_pipe_nodes(refine: int = 1):
    """(xnodes, ynodes) of the crooked-pipe mesh; refine = 1 reproduces the shipped 107 x 48 nodes (to the deck's
    5-6 printed digits), refine = k splits every shipped cell into k equal parts (interfaces stay on nodes)."""
    unit = _graded_unit()

    def axis(length, refine_map, digits):
        nodes = set(np.round(np.arange(0, int(round(length * 10)) + 1) * 0.1, 10))
        for edge, side in refine_map.items():
            u = np.round(unit, digits.get(edge, 5))  # the deck prints these offsets with 5 or 6 decimals
            if side == "left":
                nodes.update(np.round(edge - u, 10))
            else:
                nodes.update(np.round(edge + u, 10))
        return np.array(sorted(nodes))

    xn, yn = axis(7.0, CP_X_REFINE, {}), axis(2.0, CP_Y_REFINE, {0.5: 6, 1.0: 6})
    if refine > 1:
        def split(n):
            parts = [np.linspace(n[i], n[i + 1], refine, endpoint=False) for i in range(len(n) - 1)]
            return np.concatenate(parts + [[n[-1]]])
        xn, yn = split(xn), split(yn)
    return xn, yn


def crooked_pipe_nodes_sized(nx_cells: int, ny_cells: int):
    """Crooked-pipe nodes resampled to exactly (nx_cells, ny_cells) cells (BASELINE configs 3 and 5: 1024^2, 4096^2):
    every shipped cell is split into equal parts, the number of parts chosen per cell so the totals match; the
    material interfaces (shipped nodes) all remain nodes."""
    bx, by = crooked_pipe_nodes(1)

    def resample(base, n_cells):
        nb = len(base) - 1
        if n_cells < nb:
            raise ValueError(f"need at least {nb} cells")
        q, r = divmod(n_cells, nb)
        # distribute the remainder to the widest cells first (deterministic)
        widths = np.diff(base)
        order = np.argsort(-widths, kind="stable")
        parts = np.full(nb, q, dtype=np.int64)
        parts[order[:r]] += 1
        out = [np.linspace(base[i], base[i + 1], parts[i], endpoint=False) for i in range(nb)]
        return np.concatenate(out + [[base[-1]]])

    return resample(bx, nx_cells), resample(by, ny_cells)


CP_RECTS = (((0.0, 7.0), (0.0, 2.0)), ((0.0, 3.0), (0.0, 0.5)), ((2.5, 3.0), (0.0, 1.5)), ((2.5, 4.5), (1.0, 1.5)),
            ((4.0, 4.5), (0.0, 1.5)), ((4.0, 7.0), (0.0, 0.5)))


def crooked_pipe(precision="FLOAT64", n_input=50000, n_max=60000, cellmin=10, pairwise="TRUE", seed=12345,
                 mesh_cells: Optional[Sequence[int]] = None, refine: int = 1, energyscales=(1.0,)) -> Dict:
    """Crooked pipe (src/inputs/CrookedPipe.txt): thick wall sigma_a 2000 / cv 1.0, pipe 0.2 / 1e-3, left surface
    T = 0.5 for y < 0.5, L/R/T VACUUM, B REFLECT.  mesh_cells=(1024, 1024) etc. gives the scaled configs."""
    T = _T(precision)
    d = _common("Crooked Pipe", precision, seed, "2D")
    if mesh_cells is not None:
        xn, yn = crooked_pipe_nodes_sized(int(mesh_cells[0]), int(mesh_cells[1]))
    else:
        xn, yn = crooked_pipe_nodes(refine)
    d.update({
        "MESHTYPE": "NONUNIFORM", "XMESHNODES": xn, "YMESHNODES": yn,
        "SIGMA_A_REGS": [CP_RECTS], "SIGMA_A_VALS": _arr(T, [2000.0, 0.2, 0.2, 0.2, 0.2, 0.2]),
        "SIGMA_A_POWERS": _arr(T, [0.0] * 6),
        "SIGMA_S_REGS": [""], "SIGMA_S_VALS": ["0.0"], "SIGMA_S_POWERS": ["0.0"],
        "BEE_REGS": [CP_RECTS], "BEE_VALS": _arr(T, [1.0, 0.001, 0.001, 0.001, 0.001, 0.001]),
        "RADSOURCE_REGS": [""], "RADSOURCE_VALS": ["0"],
        "LEFTBC": "VACUUM", "RIGHTBC": "VACUUM", "TOPBC": "VACUUM", "BOTTOMBC": "REFLECT",
        "TIMESTEPPING": "RAMP", "DT0": "0.001", "K": "1.1", "DTMAX": "0.1", "ENDTIME": "10.0",
        "NINPUT": str(n_input), "NMAX": str(n_max), "CELLMIN": str(cellmin),
        "T_INIT": "0.05",
        "T_SURFACE_VALS": [(0.0, 0.0, (0.5, 0.0), 0.0)], "T_SURFACE_REGS": [(7.0, 7.0, (0.5, 2.0), 2.0)],
        "PHYS_C": "299.70", "PHYS_A": "0.01372016", "ALPHA": "1.0",
        "LINEARIZED": "FALSE", "PAIRWISE": pairwise, "RANDOMWALK": "FALSE",
        "ENERGYSCALES": _scales(T, list(energyscales)), "DISTANCESCALE": "1.0",
    })
    return d


def small_2d(precision="FLOAT64", n_input=500, n_max=100000, bcs=("REFLECT", "VACUUM", "REFLECT", "REFLECT"), seed=12345,
             pairwise="FALSE", energyscales=(1.0,)) -> Dict:
    """The reference's own test deck (test/test_input.txt): 2-D NONUNIFORM 10 x 4, sigma_a 1, radsource 1, T = 1,
    all four surfaces at T = 1, LEFT REFLECT / RIGHT VACUUM / TOP REFLECT / BOTTOM REFLECT."""
    T = _T(precision)
    d = _common("test_calc", precision, seed, "2D")
    d.update({
        "MESHTYPE": "NONUNIFORM", "XMESHNODES": np.linspace(0.0, 1.0, 11).round(10), "YMESHNODES": np.array([0.0, 0.5, 1.0, 1.5, 2.0]),
        "SIGMA_A_REGS": [""], "SIGMA_A_VALS": ["1.0"], "SIGMA_A_POWERS": ["0.0"],
        "SIGMA_S_REGS": [""], "SIGMA_S_VALS": ["0.0"], "SIGMA_S_POWERS": ["0.0"],
        "BEE_REGS": [""], "BEE_VALS": ["1.0"], "RADSOURCE_REGS": [""], "RADSOURCE_VALS": ["1.0"],
        "LEFTBC": bcs[0], "RIGHTBC": bcs[1], "TOPBC": bcs[2], "BOTTOMBC": bcs[3],
        "TIMESTEPPING": "CONSTANT", "DT": "0.01", "ENDTIME": "1.0",
        "NINPUT": str(n_input), "NMAX": str(n_max), "CELLMIN": "1",
        "T_INIT": "1.0", "T_SURFACE_REGS": [(1.0, 1.0, 2.0, 2.0)], "T_SURFACE_VALS": [(1.0, 1.0, 1.0, 1.0)],
        "PHYS_C": "299.70", "PHYS_A": "0.01372016", "ALPHA": "1.0",
        "LINEARIZED": "FALSE", "PAIRWISE": pairwise, "RANDOMWALK": "FALSE",
        "ENERGYSCALES": _scales(T, list(energyscales)), "DISTANCESCALE": "1.0",
    })
    return d


def nonuniform_1d(precision="FLOAT64", n_input=10000, n_max=50000, seed=12345,
                  energyscales=(32768.0, 8192.0, 4096.0, 1024.0, 256.0, 64.0, 4.0, 1.0, 0.5), pairwise="TRUE") -> Dict:
    """src/inputs/1DNonUniform.txt: the crooked-pipe x nodes as a 1-D mesh with nine energy scales."""
    T = _T(precision)
    d = _common("1D NONUNIFORM", precision, seed, "1D")
    d.update({
        "MESHTYPE": "NONUNIFORM", "MESHNODES": crooked_pipe_nodes(1)[0],
        "SIGMA_A_REGS": [""], "SIGMA_A_VALS": ["0.2"], "SIGMA_A_POWERS": ["0.0"],
        "SIGMA_S_REGS": [""], "SIGMA_S_VALS": ["0.0"], "SIGMA_S_POWERS": ["0.0"],
        "BEE_REGS": [""], "BEE_VALS": ["0.001"], "RADSOURCE_REGS": [""], "RADSOURCE_VALS": ["0"],
        "LEFTBC": "VACUUM", "RIGHTBC": "VACUUM",
        "TIMESTEPPING": "CONSTANT", "DT": "0.001", "ENDTIME": "0.05",
        "NINPUT": str(n_input), "NMAX": str(n_max), "CELLMIN": "1",
        "T_INIT": "0.05", "T_SURFACE_VALS": _arr(T, [0.5, 0.0]), "T_SURFACE_REGS": [""],
        "PHYS_C": "299.70", "PHYS_A": "0.01372016", "ALPHA": "1.0",
        "LINEARIZED": "FALSE", "PAIRWISE": pairwise, "RANDOMWALK": "FALSE",
        "ENERGYSCALES": _scales(T, list(energyscales)), "DISTANCESCALE": "1.0",
    })
    return d
