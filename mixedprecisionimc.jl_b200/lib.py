"""ctypes binding of the C ABI in include/imc.h.

The product library is ``libimc_b200.so`` (CUDA, sm_100a), built in-tree by ``__graft_entry__.build()``.
There is NO CPU fallback: if the CUDA library is missing or no GPU is present, creating an engine
fails loudly.  Tests load the CPU oracle (oracle/_build/libimc_oracle.so) through the same binding by
passing its path explicitly — it exports the identical ABI.

This file is the Python twin of the Julia ``ccall`` stubs shown in INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.path.join(HERE, "libimc_b200.so")

IMC_MAX_SCALES = 16
F16, F32, F64 = 0, 1, 2
REFLECT, VACUUM = 0, 1
RNG_PHILOX, RNG_TAPE = 0, 1
TALLY_AUTO, TALLY_ATOMIC, TALLY_FIXED, TALLY_EXACT = 0, 1, 2, 3
TRACK_AUTO, TRACK_HISTORY, TRACK_REFILL, TRACK_EVENT = 0, 1, 2, 3
BC_LEFT, BC_RIGHT, BC_TOP, BC_BOTTOM = 0, 1, 2, 3

PRECISION_DTYPES = {F16: np.float16, F32: np.float32, F64: np.float64}
PRECISION_IDS = {np.dtype(np.float16): F16, np.dtype(np.float32): F32, np.dtype(np.float64): F64}

FIELDS = {
    "temp": 0, "fleck": 1, "beta": 2, "bee": 3, "sigma_a": 4, "sigma_s": 5, "energydep": 6,
    "emittedenergy": 7, "matenergydens": 8, "radenergydens": 9, "nrg_inc": 10,
}
_MULTISCALE = {"energydep", "emittedenergy"}


class ImcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"imc error {code}: {msg}")
        self.code = code


class _Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("precision", C.c_int32), ("geometry", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("bc", C.c_int32 * 4),
        ("linearized", C.c_int32), ("pairwise", C.c_int32), ("randomwalk", C.c_int32),
        ("marshak_quirk", C.c_int32), ("n_scales", C.c_int32),
        ("energyscales", C.c_double * IMC_MAX_SCALES),
        ("distancescale", C.c_double), ("phys_c", C.c_double), ("phys_a", C.c_double), ("alpha", C.c_double),
        ("seed", C.c_int64), ("n_max", C.c_int64),
        ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("rng_mode", C.c_int32), ("tally_mode", C.c_int32), ("track_mode", C.c_int32),
        ("exact_record_budget", C.c_int64),
    ]


class _SourceStats(C.Structure):
    _fields_ = [("totalenergy", C.c_double), ("emitted_sum", C.c_double), ("n_source", C.c_int64),
                ("n_new_global", C.c_int64), ("n_new_local", C.c_int64), ("n_particles", C.c_int64)]


class _TransportStats(C.Structure):
    _fields_ = [("lostenergy", C.c_double), ("segments", C.c_uint64), ("segments_total", C.c_uint64),
                ("histories", C.c_int64), ("n_census", C.c_int64), ("n_absorbed", C.c_int64),
                ("n_escaped", C.c_int64), ("n_rw", C.c_int64), ("n_errors", C.c_int64),
                ("variant", C.c_int32), ("tally_mode", C.c_int32), ("kernel_ms", C.c_float)]


class _TallyStats(C.Structure):
    _fields_ = [("totalenergydep", C.c_double), ("energy_increase", C.c_double), ("max_temp", C.c_double),
                ("total_energy_density", C.c_double)]


class _EnergyStats(C.Structure):
    _fields_ = [("radenergy", C.c_double), ("radenergy_change", C.c_double), ("lostenergy", C.c_double),
                ("energy_error", C.c_double)]


def _as_dict(s: C.Structure) -> dict:
    return {n: getattr(s, n) for n, _ in s._fields_}


_DP = C.POINTER(C.c_double)

# every symbol include/imc.h declares: (name, restype, argtypes)
ABI = [
    ("imc_abi_version", C.c_int, []),
    ("imc_backend", C.c_char_p, []),
    ("imc_create", C.c_int, [C.POINTER(_Config), C.POINTER(C.c_void_p)]),
    ("imc_destroy", None, [C.c_void_p]),
    ("imc_last_error", C.c_char_p, [C.c_void_p]),
    ("imc_set_mesh", C.c_int, [C.c_void_p] + [_DP] * 14),
    ("imc_rw_table", C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int32, _DP, _DP, _DP]),
    ("imc_update", C.c_int, [C.c_void_p, C.c_double]),
    ("imc_source", C.c_int, [C.c_void_p, C.c_double, C.c_int64, C.c_double, C.c_int64, C.c_int64, C.POINTER(_SourceStats)]),
    ("imc_transport", C.c_int, [C.c_void_p, C.c_double, C.c_int64, C.POINTER(_TransportStats)]),
    ("imc_clean", C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    ("imc_tally", C.c_int, [C.c_void_p, C.c_double, C.c_double, C.POINTER(_TallyStats)]),
    ("imc_tally_local", C.c_int, [C.c_void_p]),
    ("imc_tally_finish", C.c_int, [C.c_void_p, C.c_double, C.c_double, C.POINTER(_TallyStats)]),
    ("imc_energycheck", C.c_int, [C.c_void_p, C.POINTER(_EnergyStats)]),
    ("imc_step", C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_double, C.c_int64,
                           C.POINTER(_SourceStats), C.POINTER(_TransportStats), C.POINTER(_TallyStats), C.POINTER(_EnergyStats)]),
    ("imc_reduce_buffer", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    ("imc_get_field", C.c_int, [C.c_void_p, C.c_int32, _DP, C.c_int64]),
    ("imc_set_state", C.c_int, [C.c_void_p, _DP, _DP, _DP]),
    ("imc_field_elsize", C.c_int32, [C.c_void_p, C.c_int32]),
    ("imc_get_field_native", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]),
    ("imc_set_state_native", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("imc_stream", C.c_void_p, [C.c_void_p]),
    ("imc_history_enable", C.c_int, [C.c_void_p, C.c_int64]),
    ("imc_history_count", C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("imc_history_get", C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]),
    ("imc_history_clear", C.c_int, [C.c_void_p]),
    ("imc_num_particles", C.c_int64, [C.c_void_p]),
    ("imc_kernel_launches", C.c_int64, [C.c_void_p]),
    ("imc_get_particles", C.c_int, [C.c_void_p, _DP, C.POINTER(C.c_uint64), C.c_int64]),
    ("imc_set_particles", C.c_int, [C.c_void_p, _DP, C.POINTER(C.c_uint64), C.c_int64]),
    ("imc_set_transport_tape", C.c_int, [C.c_void_p, _DP, C.c_int32, _DP, C.c_int32, C.c_int64]),
    ("imc_set_source_tape", C.c_int, [C.c_void_p, _DP, C.c_int32, C.c_int64]),
    ("imc_get_outcomes", C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64]),
    ("imc_sample_planck", C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_double)]),
    ("imc_checkpoint", C.c_int, [C.c_void_p, C.c_int32]),
]


class ImcLib:
    """A loaded shared library exporting include/imc.h."""

    def __init__(self, path: Optional[str] = None):
        self.path = path or CUDA_LIB_PATH
        if not os.path.exists(self.path):
            raise FileNotFoundError(
                f"{self.path} not found — build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback for the transport step.")
        self.dll = C.CDLL(self.path)
        for name, res, args in ABI:
            fn = getattr(self.dll, name)  # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        ver = self.dll.imc_abi_version()
        if ver != 1:
            raise RuntimeError(f"{self.path}: ABI version {ver}, expected 1")
        self.backend = self.dll.imc_backend().decode()


_cuda_lib: Optional[ImcLib] = None


def cuda_lib() -> ImcLib:
    global _cuda_lib
    if _cuda_lib is None:
        _cuda_lib = ImcLib(CUDA_LIB_PATH)
    return _cuda_lib


@dataclass
class Config:
    """Mirror of imc_config (include/imc.h)."""
    precision: int = F64
    geometry: int = 1
    nx: int = 1
    ny: int = 1
    bc: Sequence[int] = (REFLECT, VACUUM, VACUUM, VACUUM)  # left, right, top, bottom
    linearized: bool = False
    pairwise: bool = False
    randomwalk: bool = False
    marshak_quirk: bool = False
    energyscales: Sequence[float] = (1.0,)
    distancescale: float = 1.0
    phys_c: float = 1.0
    phys_a: float = 1.0
    alpha: float = 1.0
    seed: int = 0
    n_max: int = 1 << 40
    device: int = 0
    rank: int = 0
    world: int = 1
    rng_mode: int = RNG_PHILOX
    tally_mode: int = TALLY_AUTO
    track_mode: int = TRACK_AUTO
    exact_record_budget: int = 0

    def to_c(self) -> _Config:
        c = _Config()
        c.struct_size = C.sizeof(_Config)
        c.precision, c.geometry, c.nx, c.ny = self.precision, self.geometry, self.nx, self.ny
        bc = list(self.bc) + [VACUUM] * (4 - len(self.bc))
        for i in range(4):
            c.bc[i] = int(bc[i])
        c.linearized, c.pairwise = int(self.linearized), int(self.pairwise)
        c.randomwalk, c.marshak_quirk = int(self.randomwalk), int(self.marshak_quirk)
        scales = sorted((float(s) for s in self.energyscales), reverse=True)
        c.n_scales = len(scales)
        for i, s in enumerate(scales):
            c.energyscales[i] = s
        c.distancescale, c.phys_c, c.phys_a, c.alpha = self.distancescale, self.phys_c, self.phys_a, self.alpha
        c.seed, c.n_max = int(self.seed), int(self.n_max)
        c.device, c.rank, c.world = self.device, self.rank, self.world
        c.rng_mode, c.tally_mode, c.track_mode = self.rng_mode, self.tally_mode, self.track_mode
        c.exact_record_budget = self.exact_record_budget
        return c


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data_as(_DP)


def _f64(a, n: Optional[int] = None) -> np.ndarray:
    """Contiguous Float64 copy in Julia (column-major) linear order."""
    out = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, order="F"))
    if n is not None and out.size != n:
        raise ValueError(f"expected {n} values, got {out.size}")
    return out


class Engine:
    """One transport-step engine (one GPU).  Thin, explicit wrapper over the C ABI."""

    def __init__(self, cfg: Config, lib: Optional[ImcLib] = None):
        self.lib = lib or cuda_lib()
        self.cfg = cfg
        self._h = C.c_void_p()
        self.nc = cfg.nx * (cfg.ny if cfg.geometry == 2 else 1)
        self.ns = len(cfg.energyscales)
        self.nslots = 9 if cfg.geometry == 1 else 10
        ccfg = cfg.to_c()
        rc = self.lib.dll.imc_create(C.byref(ccfg), C.byref(self._h))
        if rc != 0:
            raise ImcError(rc, (self.lib.dll.imc_last_error(None) or b"").decode())

    # -- lifetime --------------------------------------------------------------------------
    def close(self):
        if self._h:
            self.lib.dll.imc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise ImcError(rc, (self.lib.dll.imc_last_error(self._h) or b"").decode())

    # -- setup -----------------------------------------------------------------------------
    def set_mesh(self, *, dx, dy=None, sigma_a_const, sigma_a_pow, sigma_s_const, sigma_s_pow, sigma_static=None,
                 bee, radsource, temp, tsurf_bottom=None, tsurf_top=None, tsurf_left=None, tsurf_right=None):
        c = self.cfg
        nc = self.nc
        keep = [  # keep arrays alive for the duration of the call
            _f64(dx, c.nx), _f64(dy, c.ny) if c.geometry == 2 else None,
            _f64(sigma_a_const, nc), _f64(sigma_a_pow, nc), _f64(sigma_s_const, nc), _f64(sigma_s_pow, nc),
            _f64(sigma_static, nc) if sigma_static is not None else None,
            _f64(bee, nc), _f64(radsource, nc), _f64(temp, nc),
            _f64(tsurf_bottom, c.nx) if c.geometry == 2 else None,
            _f64(tsurf_top, c.nx) if c.geometry == 2 else None,
            _f64(tsurf_left, c.ny if c.geometry == 2 else 1),
            _f64(tsurf_right, c.ny if c.geometry == 2 else 1),
        ]
        self._check(self.lib.dll.imc_set_mesh(self._h, *[_dp(a) for a in keep]))

    def rw_table(self, a_lo=0.0, a_hi=10.0, n=1000):
        a = np.empty(n); pr = np.empty(n); pt = np.empty(n)
        self._check(self.lib.dll.imc_rw_table(self._h, a_lo, a_hi, n, _dp(a), _dp(pr), _dp(pt)))
        return a, pr, pt

    # -- the per-stage calls (reference call sites, MixedPrecisionIMC.jl:138-147) ---------------
    def update(self, dt: float):
        self._check(self.lib.dll.imc_update(self._h, float(dt)))

    def source(self, dt: float, n_input: int, cellmin: float, step: int, n_census_global: int = -1) -> dict:
        st = _SourceStats()
        self._check(self.lib.dll.imc_source(self._h, float(dt), int(n_input), float(cellmin), int(step), int(n_census_global), C.byref(st)))
        return _as_dict(st)

    def transport(self, dt: float, step: int) -> dict:
        st = _TransportStats()
        self._check(self.lib.dll.imc_transport(self._h, float(dt), int(step), C.byref(st)))
        return _as_dict(st)

    def clean(self) -> int:
        n = C.c_int64()
        self._check(self.lib.dll.imc_clean(self._h, C.byref(n)))
        return n.value

    def tally(self, t: float, dt: float) -> dict:
        st = _TallyStats()
        self._check(self.lib.dll.imc_tally(self._h, float(t), float(dt), C.byref(st)))
        return _as_dict(st)

    def tally_local(self):
        self._check(self.lib.dll.imc_tally_local(self._h))

    def tally_finish(self, t: float, dt: float) -> dict:
        st = _TallyStats()
        self._check(self.lib.dll.imc_tally_finish(self._h, float(t), float(dt), C.byref(st)))
        return _as_dict(st)

    def energycheck(self) -> dict:
        st = _EnergyStats()
        self._check(self.lib.dll.imc_energycheck(self._h, C.byref(st)))
        return _as_dict(st)

    def step(self, t: float, dt: float, n_input: int, cellmin: float, step: int) -> dict:
        a, b, c_, d = _SourceStats(), _TransportStats(), _TallyStats(), _EnergyStats()
        self._check(self.lib.dll.imc_step(self._h, float(t), float(dt), int(n_input), float(cellmin), int(step),
                                          C.byref(a), C.byref(b), C.byref(c_), C.byref(d)))
        return {"source": _as_dict(a), "transport": _as_dict(b), "tally": _as_dict(c_), "energy": _as_dict(d)}

    def checkpoint(self, op: str = "save"):
        """Restart point inside the library (include/imc.h imc_checkpoint): 'save' | 'restore' | 'drop'."""
        self._check(self.lib.dll.imc_checkpoint(self._h, {"save": 0, "restore": 1, "drop": 2}[op]))

    def reduce_buffer(self):
        """(address, n_elements, kind) of the buffer to all-reduce between tally_local and tally_finish; kind: 0 = Float64,
        1 = int64 (FIXED tallies) (include/imc.h)."""
        p = C.c_void_p(); n = C.c_int64(); kind = C.c_int32()
        self._check(self.lib.dll.imc_reduce_buffer(self._h, C.byref(p), C.byref(n), C.byref(kind)))
        return p.value, n.value, int(kind.value)

    # -- state access ------------------------------------------------------------------------
    def field(self, name: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Download a field as Float64; `out` may be a caller-owned (e.g. pinned) flat buffer."""
        n = self.nc * (self.ns if name in _MULTISCALE else 1)
        if out is None:
            out = np.empty(n, dtype=np.float64)
        assert out.size == n and out.dtype == np.float64 and out.flags["C_CONTIGUOUS"]
        self._check(self.lib.dll.imc_get_field(self._h, FIELDS[name], _dp(out), n))
        c = self.cfg
        if c.geometry == 2:
            shape = (c.nx, c.ny, self.ns) if name in _MULTISCALE else (c.nx, c.ny)
        else:
            shape = (c.nx, self.ns) if name in _MULTISCALE else (c.nx,)
        return out.reshape(shape, order="F")

    def set_state(self, temp=None, matenergydens=None, radenergydens=None):
        arrs = [None if a is None else _f64(a, self.nc) for a in (temp, matenergydens, radenergydens)]
        self._check(self.lib.dll.imc_set_state(self._h, *[_dp(a) for a in arrs]))

    # ---- the same transfers in the field's own element type (what the Julia shim does with pointer(mesh.x)) ----
    def field_dtype(self, name: str):
        es = int(self.lib.dll.imc_field_elsize(self._h, FIELDS[name]))
        return {2: np.float16, 4: np.float32, 8: np.float64}[es]

    def field_native(self, name: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Download a field in its own element type (Array{T}); `out` may be a caller-owned (pinned) flat buffer."""
        n = self.nc * (self.ns if name in _MULTISCALE else 1)
        dt = self.field_dtype(name)
        if out is None:
            out = np.empty(n, dtype=dt)
        assert out.size == n and out.dtype == dt and out.flags["C_CONTIGUOUS"]
        self._check(self.lib.dll.imc_get_field_native(self._h, FIELDS[name], out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def set_state_native(self, temp=None, matenergydens=None, radenergydens=None):
        """Upload fields in their own element type.  Each argument is a flat numpy array (host memory, pinned or not) or a
        contiguous torch tensor (host or CUDA memory of this process: the engine copies with unified addressing)."""
        def ptr(a, name):
            if a is None:
                return None
            if hasattr(a, "data_ptr"):      # torch tensor
                assert a.numel() == self.nc and a.element_size() == np.dtype(self.field_dtype(name)).itemsize and a.is_contiguous(), name
                return C.c_void_p(a.data_ptr())
            assert a.size == self.nc and a.dtype == self.field_dtype(name) and a.flags["C_CONTIGUOUS"], name
            return a.ctypes.data_as(C.c_void_p)
        self._check(self.lib.dll.imc_set_state_native(self._h, ptr(temp, "temp"), ptr(matenergydens, "matenergydens"),
                                                      ptr(radenergydens, "radenergydens")))

    # ---- per-step history kept by the engine (mesh.temp_saved / matenergy_saved / radenergy_saved / energyincrease_saved) ----
    def history_enable(self, capacity: int):
        self._check(self.lib.dll.imc_history_enable(self._h, int(capacity)))

    def history_count(self):
        n, d = C.c_int64(), C.c_int64()
        self._check(self.lib.dll.imc_history_count(self._h, C.byref(n), C.byref(d)))
        return n.value, d.value

    def history(self, name: str, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """[count, Nc] snapshots of `name` (temp: Float64; matenergydens / radenergydens / nrg_inc: the deck precision)."""
        if count is None:
            count = self.history_count()[0] - first
        dt = np.float64 if name == "temp" else PRECISION_DTYPES[self.cfg.precision]
        out = np.empty((count, self.nc), dtype=dt)
        self._check(self.lib.dll.imc_history_get(self._h, FIELDS[name], int(first), int(count), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def history_clear(self):
        self._check(self.lib.dll.imc_history_clear(self._h))

    def stream(self) -> int:
        """cudaStream_t of the engine as an integer (0 for the oracle)."""
        return int(self.lib.dll.imc_stream(self._h) or 0)

    def num_particles(self) -> int:
        return int(self.lib.dll.imc_num_particles(self._h))

    def kernel_launches(self) -> int:
        return int(self.lib.dll.imc_kernel_launches(self._h))

    def particles(self):
        """(slots [N, 9|10] Float64 in the reference's slot order, ids [N] uint64)."""
        n = self.num_particles()
        slots = np.empty((n, self.nslots), dtype=np.float64)
        ids = np.empty(n, dtype=np.uint64)
        self._check(self.lib.dll.imc_get_particles(self._h, _dp(slots), ids.ctypes.data_as(C.POINTER(C.c_uint64)), n))
        return slots, ids

    def set_particles(self, slots: np.ndarray, ids: Optional[np.ndarray] = None):
        slots = np.ascontiguousarray(slots, dtype=np.float64)
        assert slots.ndim == 2 and slots.shape[1] == self.nslots
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint64)
            idp = ids.ctypes.data_as(C.POINTER(C.c_uint64))
        self._check(self.lib.dll.imc_set_particles(self._h, _dp(slots), idp, slots.shape[0]))

    def set_transport_tape(self, uniforms: np.ndarray, exponentials: np.ndarray):
        """uniforms [n_uni, n_slots], exponentials [n_exp, n_slots] (draw-major)."""
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        e = np.ascontiguousarray(exponentials, dtype=np.float64)
        assert u.shape[1] == e.shape[1]
        self._check(self.lib.dll.imc_set_transport_tape(self._h, _dp(u), u.shape[0], _dp(e), e.shape[0], u.shape[1]))

    def set_source_tape(self, uniforms: np.ndarray):
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        self._check(self.lib.dll.imc_set_source_tape(self._h, _dp(u), u.shape[0], u.shape[1]))

    def sample_planck(self, n: int, step: int = 0) -> np.ndarray:
        """``Sourcing.sample_planck`` (imc_sourcing.jl:372-399): n Planck-spectrum frequencies in the deck precision."""
        out = np.empty(n, dtype=np.float64)
        self._check(self.lib.dll.imc_sample_planck(self._h, n, step, _dp(out)))
        return out

    def outcomes(self, n: int):
        ev = np.empty(n, dtype=np.int32); ns = np.empty(n, dtype=np.int32)
        self._check(self.lib.dll.imc_get_outcomes(self._h, ev.ctypes.data_as(C.POINTER(C.c_int32)),
                                                  ns.ctypes.data_as(C.POINTER(C.c_int32)), n))
        return ev, ns
