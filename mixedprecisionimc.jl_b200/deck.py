"""Host-side look-alike of the reference's deck parser and mesh generator.

In production these stay Julia (``Input.readInputs`` src/imc_input.jl:20-95, ``Constants.set_constants``
src/constants.jl:6-13, ``Mesh.mesh_generation`` src/imc_mesh.jl:42-173): they are OUT OF SCOPE for the
engine.  Julia is not available in this environment, so this module restates them in
Python/numpy to drive the engine and the oracle from the reference's own deck files in tests and
benchmarks.  Values are parsed *through the deck precision* exactly as the reference does (numpy
float16/32/64 scalars), so what crosses the C ABI is what the Julia host would pass.
"""
from __future__ import annotations

import ast
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

_PRECISIONS = {
    "HALF": np.float16, "FLOAT16": np.float16,
    "SINGLE": np.float32, "FLOAT32": np.float32,
    "DOUBLE": np.float64, "FLOAT64": np.float64,
}


def parse_T(T, s) -> Any:
    """Julia ``parse(T, str)``; Float16 parses through Float32 (base/parse.jl)."""
    if isinstance(s, (bytes, str)):
        s = s.strip()
    if T is np.float16:
        with np.errstate(over="ignore"):   # parse(Float16, "100000") is Inf16 in Julia too
            return np.float16(np.float32(s))
    return T(s)


def read_inputs(filename: str) -> Dict[str, Any]:
    """``Input.readInputs`` (imc_input.jl:20-95): ``key = value`` lines, ``[a, b]`` arrays that may span
    lines, ``#`` comments.  Arrays with commas are parsed (mesh nodes always as Float64, :85-86); arrays
    containing ``(`` are evaluated as nested tuples (:105-106); single-element arrays stay strings (Q22)."""
    params: Dict[str, Any] = {}
    key, cur, in_array = "", [], False
    with open(filename, "r") as f:
        for raw in f:
            line = raw.strip()
            if not line or line.startswith("#"):
                continue
            if not in_array and re.match(r"^\w+ = ", line):
                parts = line.split("=")
                key = parts[0].strip()
                val = parts[1].strip()
                if not val.startswith("["):
                    params[key] = val
                else:
                    if not val.endswith("]"):
                        in_array = True
                    cur = [val.strip("[]")]
                    if not in_array:
                        params[key] = cur
            elif in_array:
                if line.endswith("]"):
                    cur.append(line.strip("]"))
                    params[key] = cur
                    in_array = False
                else:
                    cur.append(line)
    if "PRECISION" in params and params["PRECISION"] in _PRECISIONS:
        params["PRECISION"] = _PRECISIONS[params["PRECISION"]]
    T = params["PRECISION"]
    for k, v in list(params.items()):
        if k == "PRECISION":
            continue
        joined = "".join(v) if isinstance(v, list) else v
        if "," in joined:
            if k in ("MESHNODES", "XMESHNODES", "YMESHNODES"):
                params[k] = _separated_array_parser(v, np.float64)
            else:
                params[k] = _separated_array_parser(v, T)
    return params


def _separated_array_parser(value, T):
    joined = "".join(value) if isinstance(value, list) else value
    if "(" in joined:
        return [ast.literal_eval(joined.strip())]  # [eval(Meta.parse(...))]: one element holding the tuple
    return np.array([parse_T(T, s) for s in joined.split(",")], dtype=T)


@dataclass
class Constants:
    phys_c: Any
    phys_a: Any
    alpha: Any


def set_constants(inputs) -> Constants:
    T = inputs["PRECISION"]
    return Constants(parse_T(T, inputs["PHYS_C"]), parse_T(T, inputs["PHYS_A"]), parse_T(T, inputs["ALPHA"]))


def tointeger(x) -> int:
    """``Utilities.tointeger`` (imc_utilities.jl:6-21): exact conversion or InexactError."""
    xf = float(x)
    if not np.isfinite(xf) or xf != int(xf):
        raise ValueError(f"InexactError: {x!r} is not an integer")
    return int(xf)


def _scalar(v):
    """Single-element arrays stay as a 1-list of strings in the reference; unwrap."""
    if isinstance(v, list) and len(v) == 1 and isinstance(v[0], str):
        return v[0]
    return v


@dataclass
class MeshStruct:
    """Subset of ``MeshStruct`` (imc_mesh.jl:10-40) that the transport step reads or writes."""
    geometry: str
    precision: Any
    Ncells: Any                 # int or (Nx, Ny)
    dx: np.ndarray
    dy: Any
    nodes: Any
    centers: Any
    temp: np.ndarray
    temp_surf: Any
    sigma_a: np.ndarray         # [..., 3]: current / constant / power
    sigma_s: np.ndarray
    sigma: np.ndarray           # sigma_a + sigma_s (static, Q4)
    bee: np.ndarray
    radsource: np.ndarray
    energyscales: Any
    distancescale: Any
    # tallies and diagnostics mirrored back from the engine after each stage
    fleck: Optional[np.ndarray] = None
    beta: Optional[np.ndarray] = None
    energydep: Optional[np.ndarray] = None
    emittedenergy: Optional[np.ndarray] = None
    radenergydens: Optional[np.ndarray] = None
    matenergydens: Optional[np.ndarray] = None
    totalenergy: float = 0.0
    totalenergydep: float = 0.0
    lostenergy: float = 0.0
    radenergyold: float = 0.0
    temp_saved: List[np.ndarray] = field(default_factory=list)
    radenergy_saved: List[np.ndarray] = field(default_factory=list)
    matenergy_saved: List[np.ndarray] = field(default_factory=list)
    energyincrease_saved: List[np.ndarray] = field(default_factory=list)
    engine: Any = None          # the imc engine that owns the device copy

    @property
    def nx(self) -> int:
        return self.Ncells if self.geometry == "1D" else self.Ncells[0]

    @property
    def ny(self) -> int:
        return 1 if self.geometry == "1D" else self.Ncells[1]


def _linrange(T, a, b, n):
    """``LinRange(a, b, n)`` with element type T: lerp evaluated in Float64 then converted (base/range.jl lerpi)."""
    if n == 1:
        return np.array([a], dtype=T)
    t = np.arange(n, dtype=np.float64) / float(n - 1)
    return ((1.0 - t) * float(a) + t * float(b)).astype(T)


def region_joiner(geometry, regions, values, nodes, Ncells, T):
    """``Mesh.region_joiner`` (imc_mesh.jl:175-220)."""
    regions = _scalar(regions)
    if geometry == "1D":
        nreg = len(regions) if isinstance(regions, (list, np.ndarray, tuple)) else 1
        if nreg <= 1:
            return np.full(Ncells, parse_T(T, _first(values)), dtype=T)
        out = np.zeros(Ncells, dtype=T)
        ri = 0
        for i in range(Ncells):
            if not (regions[ri] >= nodes[i + 1]):
                ri += 1
            out[i] = values[ri]
        return out
    # 2-D: regions is [tuple_of_rectangles] or [] / one string
    rects = regions[0] if isinstance(regions, list) and len(regions) == 1 and isinstance(regions[0], tuple) else None
    if rects is None or len(rects) <= 1:
        return np.full(Ncells, parse_T(T, _first(values)), dtype=T)
    out = np.zeros(Ncells, dtype=T)
    xn, yn = np.asarray(nodes[0], dtype=np.float64), np.asarray(nodes[1], dtype=np.float64)
    for ri, rect in enumerate(rects):
        (xs, xe), (ys, ye) = rect
        xsi = _findlast_le(xn, xs); xei = _findlast_le(xn, xe) - 1
        ysi = _findlast_le(yn, ys); yei = _findlast_le(yn, ye) - 1
        out[xsi - 1:xei, ysi - 1:yei] = values[ri]  # Julia 1-based inclusive ranges
    return out


def _first(values):
    v = _scalar(values)
    if isinstance(v, (list, tuple, np.ndarray)):
        return v[0]
    return v


def _findlast_le(arr, x) -> int:
    idx = np.nonzero(arr <= x)[0]
    return int(idx[-1]) + 1  # 1-based


def surface_definer(geometry, regions, values, nodes, Ncells, T):
    """``Mesh.surface_definer`` (imc_mesh.jl:222-259)."""
    if geometry == "1D":
        return (values[0], values[1])
    vals = values[0]
    regs = regions[0]
    out = []
    for i in range(4):  # bottom, top, left, right
        v = vals[i]
        r = regs[i]
        n1 = Ncells[0] if i < 2 else Ncells[1]
        nd = nodes[0] if i < 2 else nodes[1]
        if not isinstance(v, tuple):
            out.append(np.full(n1, parse_T(T, str(v)), dtype=T))
        else:
            rr = list(r) if isinstance(r, tuple) else [r]
            out.append(region_joiner("1D", rr, list(v), nd, n1, T))
    return out


def mesh_generation(inputs) -> MeshStruct:
    """``Mesh.mesh_generation`` (imc_mesh.jl:42-173)."""
    geometry = inputs["GEOMETRY"].upper()
    meshtype = inputs["MESHTYPE"].upper()
    T = inputs["PRECISION"]
    if geometry == "1D":
        if meshtype == "UNIFORM":
            xsize = parse_T(T, inputs["XSIZE"]); dx_t = parse_T(T, inputs["DX"])
            Ncells = tointeger(xsize / dx_t)
            dx = np.full(Ncells, dx_t, dtype=T)
            nodes = _linrange(T, 0, xsize, Ncells + 1)
            centers = _linrange(T, dx_t / T(2), xsize - dx_t / T(2), Ncells)
        else:
            n64 = np.asarray(inputs["MESHNODES"], dtype=np.float64)
            centers = ((n64[:-1] + n64[1:]) / 2).astype(T)
            dx = (n64[1:] - n64[:-1]).astype(T)  # difference in Float64, then T (imc_mesh.jl:70)
            nodes = n64.astype(T)
            Ncells = len(dx)
        dy = T(1)
    else:
        if meshtype == "UNIFORM":
            xsize = parse_T(T, inputs["XSIZE"]); ysize = parse_T(T, inputs["YSIZE"])
            dx_t = parse_T(T, inputs["DX"]); dy_t = parse_T(T, inputs["DY"])
            Ncells = (tointeger(xsize / dx_t), tointeger(ysize / dx_t))  # Q8: Ny from ysize/dx
            dx = np.full(Ncells[0], dx_t, dtype=T); dy = np.full(Ncells[1], dy_t, dtype=T)
            nodes = (_linrange(T, 0, xsize, Ncells[0] + 1), _linrange(T, 0, ysize, Ncells[1] + 1))
            centers = (_linrange(T, dx_t / T(2), xsize - dx_t / T(2), Ncells[0]),
                       _linrange(T, dy_t / T(2), ysize - dy_t / T(2), Ncells[1]))
        else:
            xn = np.asarray(inputs["XMESHNODES"], dtype=np.float64); yn = np.asarray(inputs["YMESHNODES"], dtype=np.float64)
            centers = (((xn[:-1] + xn[1:]) / 2).astype(T), ((yn[:-1] + yn[1:]) / 2).astype(T))
            dx = (xn[1:] - xn[:-1]).astype(T); dy = (yn[1:] - yn[:-1]).astype(T)
            nodes = (xn.astype(T), yn.astype(T))
            Ncells = (len(dx), len(dy))
    T_init = parse_T(T, inputs["T_INIT"])
    T_surface = surface_definer(geometry, inputs["T_SURFACE_REGS"], inputs["T_SURFACE_VALS"], nodes, Ncells, T)
    temp = np.full(Ncells, T_init, dtype=T)
    rj = lambda regs, vals: region_joiner(geometry, inputs[regs], inputs[vals], nodes, Ncells, T)
    sa = rj("SIGMA_A_REGS", "SIGMA_A_VALS"); sap = rj("SIGMA_A_REGS", "SIGMA_A_POWERS")
    ss = rj("SIGMA_S_REGS", "SIGMA_S_VALS"); ssp = rj("SIGMA_S_REGS", "SIGMA_S_POWERS")
    es = inputs["ENERGYSCALES"]
    if isinstance(es, np.ndarray):
        energyscales = np.sort(es)[::-1].astype(T)
    else:
        energyscales = np.array([parse_T(T, _first(es))], dtype=T)  # scalar in the reference (Q28)
    ds = parse_T(T, inputs["DISTANCESCALE"])
    sa = (sa / ds).astype(T); ss = (ss / ds).astype(T)
    sigma_a = np.stack([sa, sa, sap], axis=-1)
    sigma_s = np.stack([ss, ss, ssp], axis=-1)
    sigma = (sigma_a + sigma_s).astype(T)
    radsource = rj("RADSOURCE_REGS", "RADSOURCE_VALS")
    bee = rj("BEE_REGS", "BEE_VALS")
    m = MeshStruct(geometry=geometry, precision=T, Ncells=Ncells, dx=dx, dy=dy, nodes=nodes, centers=centers,
                   temp=temp, temp_surf=T_surface, sigma_a=sigma_a, sigma_s=sigma_s, sigma=sigma, bee=bee,
                   radsource=radsource, energyscales=energyscales, distancescale=ds)
    m.temp_saved.append(temp.copy())
    return m
