"""Host-side mirror of the reference's time-step driver and of the per-stage call sites.

``main(["deck.txt"])`` restates ``MixedPrecisionIMC.main`` (src/MixedPrecisionIMC.jl:59-179) and
``timestep`` (:181-222).  The namespaces ``Update``, ``Sourcing``, ``Transport``, ``Clean``, ``Tally`` and
``EnergyCheck`` keep the reference's function names and argument lists
(``Update.update(inputs, mesh, simvars)`` ... ) and forward to the engine through the C ABI — exactly
what the Julia shim in ``julia/MixedPrecisionIMCB200.jl`` does with ``ccall``.  The
engine owns the particle population, so ``particles`` is an opaque ``ParticleList`` whose ``len()`` is
served by the engine.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional

import numpy as np

from . import deck as _deck
from . import lib as _lib
from .deck import MeshStruct, parse_T, tointeger


def parse_count(T, text) -> int:
    """NINPUT / NMAX.  The reference parses them through the deck precision (MixedPrecisionIMC.jl:107-110), so a Float16
    deck cannot say more than 65504 (Q10).  The engine keeps counts as integers; a count a Float16 deck cannot represent
    is therefore taken as the integer written in the deck — the intentional divergence BASELINE configs 4-5 need
     — while every representable count goes through T exactly like the reference."""
    v = parse_T(T, text)
    if np.dtype(T) == np.dtype(np.float16) and not np.isfinite(float(v)):
        x = float(text)
        if x != int(x):
            raise ValueError(f"InexactError: {text!r} is not an integer")
        return int(x)
    return tointeger(v)


@dataclass
class SimVars:
    """``SimVars`` (MixedPrecisionIMC.jl:35-51)."""
    t: Any
    dt: Any
    dt0: Any
    k: Any
    dtmax: Any
    t_end: Any
    timesteps: List[Any]
    iterations: int
    n_input: int
    n_max: int
    cellmin: Any
    pairwise: str
    BC: tuple
    precision: Any
    geometry: str
    step: int = 0  # engine addition: time-step ordinal (Philox counter word)


@dataclass
class RWVars:
    """``RWVars`` (MixedPrecisionIMC.jl:53-57)."""
    aVals: np.ndarray
    prVals: np.ndarray
    ptVals: np.ndarray


class ParticleList:
    """Opaque stand-in for the reference's ``particles`` vector; the engine owns the data."""

    def __init__(self, engine: _lib.Engine):
        self.engine = engine

    def __len__(self) -> int:
        return self.engine.num_particles()

    def slots(self) -> np.ndarray:
        return self.engine.particles()[0]


_BC = {"REFLECT": _lib.REFLECT, "VACUUM": _lib.VACUUM}


def make_config(inputs, mesh: MeshStruct, **overrides) -> _lib.Config:
    """Everything the engine needs from the deck (SURVEY.md §5 'Config / flags')."""
    T = inputs["PRECISION"]
    geom = 1 if mesh.geometry == "1D" else 2
    consts = _deck.set_constants(inputs)
    bc = [_lib.VACUUM] * 4
    names = ["LEFTBC", "RIGHTBC", "TOPBC", "BOTTOMBC"][: 2 if geom == 1 else 4]
    for i, nm in enumerate(names):
        s = str(inputs[nm]).upper()
        if s not in _BC:
            raise ValueError(f"{nm} = {s}: the transport loop only handles REFLECT and VACUUM")
        bc[i] = _BC[s]
    n_max = parse_count(T, inputs["NMAX"])
    cfg = _lib.Config(
        precision=_lib.PRECISION_IDS[np.dtype(T)], geometry=geom, nx=mesh.nx, ny=mesh.ny, bc=bc,
        linearized=str(inputs["LINEARIZED"]).upper() == "TRUE",
        pairwise=str(inputs["PAIRWISE"]).upper() == "TRUE",
        randomwalk=geom == 1 and str(inputs.get("RANDOMWALK", "FALSE")).upper() == "TRUE",
        marshak_quirk=str(inputs["NAME"]).upper() == "MARSHAK WAVE",
        energyscales=[float(s) for s in np.atleast_1d(mesh.energyscales)],
        distancescale=float(mesh.distancescale), phys_c=float(consts.phys_c), phys_a=float(consts.phys_a),
        alpha=float(consts.alpha), seed=int(inputs["SEED"]), n_max=n_max)
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def attach_engine(inputs, mesh: MeshStruct, library: Optional[_lib.ImcLib] = None, **overrides) -> _lib.Engine:
    """Create an engine for this deck and upload the mesh (imc_create + imc_set_mesh)."""
    cfg = make_config(inputs, mesh, **overrides)
    eng = _lib.Engine(cfg, library)
    geom2 = mesh.geometry == "2D"
    ts = mesh.temp_surf
    eng.set_mesh(
        dx=mesh.dx, dy=mesh.dy if geom2 else None,
        sigma_a_const=mesh.sigma_a[..., 1], sigma_a_pow=mesh.sigma_a[..., 2],
        sigma_s_const=mesh.sigma_s[..., 1], sigma_s_pow=mesh.sigma_s[..., 2],
        sigma_static=mesh.sigma[..., 0], bee=mesh.bee, radsource=mesh.radsource, temp=mesh.temp,
        tsurf_bottom=ts[0] if geom2 else None, tsurf_top=ts[1] if geom2 else None,
        tsurf_left=ts[2] if geom2 else [ts[0]], tsurf_right=ts[3] if geom2 else [ts[1]])
    mesh.engine = eng
    return eng


# ---- the reference's stage modules, same names and argument lists ------------------------------
class Update:
    @staticmethod
    def update(inputs, mesh: MeshStruct, simvars: SimVars):
        """``Update.update`` (imc_update.jl:12)."""
        mesh.engine.update(float(simvars.dt))


class Sourcing:
    @staticmethod
    def sourcing(mesh: MeshStruct, simvars: SimVars, particles: ParticleList, n_census_global: int = -1):
        """``Sourcing.sourcing`` (imc_sourcing.jl:12)."""
        st = mesh.engine.source(float(simvars.dt), simvars.n_input, float(simvars.cellmin), simvars.step, n_census_global)
        mesh.totalenergy = st["totalenergy"]
        return st


class Transport:
    @staticmethod
    def _run(mesh, simvars):
        st = mesh.engine.transport(float(simvars.dt), simvars.step)
        simvars.iterations = st["segments_total"]
        mesh.lostenergy = st["lostenergy"]
        return st

    @staticmethod
    def MC(mesh, simvars, particles):
        """``Transport.MC`` (imc_transport.jl:13)."""
        return Transport._run(mesh, simvars)

    @staticmethod
    def MC_RW(mesh, simvars, rwvars, particles):
        """``Transport.MC_RW`` (imc_transport.jl:212)."""
        return Transport._run(mesh, simvars)

    @staticmethod
    def MC2D(mesh, simvars, particles):
        """``Transport.MC2D`` (imc_transport.jl:483)."""
        return Transport._run(mesh, simvars)

    @staticmethod
    def randomwalk_table(mesh, a_lo=0.0, a_hi=10.0, n=1000) -> RWVars:
        """``Transport.randomwalk_table`` (imc_transport.jl:786) for aVals = T.(LinRange(0, 10, 1000))."""
        return RWVars(*mesh.engine.rw_table(a_lo, a_hi, n))


class Clean:
    @staticmethod
    def clean(particles: ParticleList) -> int:
        """``Clean.clean`` (imc_clean.jl:6)."""
        return particles.engine.clean()


class Tally:
    @staticmethod
    def tally(inputs, mesh: MeshStruct, simvars: SimVars, particles: ParticleList, save: bool = True):
        """``Tally.tally`` (imc_tally.jl:11)."""
        st = mesh.engine.tally(float(simvars.t), float(simvars.dt))
        mesh.totalenergydep = st["totalenergydep"]
        if save:  # history lists (:138-142)
            mesh.temp = mesh.engine.field("temp")
            mesh.matenergydens = mesh.engine.field("matenergydens")
            mesh.radenergydens = mesh.engine.field("radenergydens")
            mesh.temp_saved.append(mesh.temp.copy())
            mesh.matenergy_saved.append(mesh.matenergydens.copy())
            mesh.radenergy_saved.append(mesh.radenergydens.copy())
        return st


class EnergyCheck:
    @staticmethod
    def energychecker(inputs, mesh: MeshStruct, simvars: SimVars, particles: ParticleList):
        """``EnergyCheck.energychecker`` (imc_energycheck.jl:10)."""
        st = mesh.engine.energycheck()
        mesh.radenergyold = st["radenergy"]
        mesh.lostenergy = 0.0
        return st


def timestep(timestepping: str, simvars: SimVars):
    """``timestep`` (MixedPrecisionIMC.jl:181-222), arithmetic in the deck precision."""
    if simvars.t == simvars.t_end:
        simvars.t = simvars.t_end + simvars.dt
        return
    if timestepping == "CONSTANT":
        if simvars.t + simvars.dt > simvars.t_end:
            simvars.dt = simvars.t_end - simvars.t
            simvars.t = simvars.t_end
        else:
            simvars.t = simvars.t + simvars.dt
    elif timestepping == "RAMP":
        if simvars.dt < simvars.dtmax:
            simvars.dt = simvars.dt * simvars.k
            if simvars.dt > simvars.dtmax:
                simvars.dt = simvars.dtmax
        if simvars.t + simvars.dt > simvars.t_end:
            simvars.dt = simvars.t_end - simvars.t
            simvars.t = simvars.t_end
        else:
            simvars.t = simvars.t + simvars.dt
    simvars.timesteps.append(simvars.t)


def make_simvars(inputs, mesh: MeshStruct) -> SimVars:
    """The SimVars construction in ``main`` (MixedPrecisionIMC.jl:88-125, :158-164)."""
    T = inputs["PRECISION"]
    ts = str(inputs["TIMESTEPPING"]).upper()
    if ts == "CONSTANT":
        dt = parse_T(T, inputs["DT"]); t_end = parse_T(T, inputs["ENDTIME"])
        dt0 = k = dtmax = T(0)
    elif ts == "RAMP":
        dt0 = parse_T(T, inputs["DT0"]); k = parse_T(T, inputs["K"]); dtmax = parse_T(T, inputs["DTMAX"])
        t_end = parse_T(T, inputs["ENDTIME"]); dt = dt0
    else:
        raise ValueError(f"TIMESTEPPING = {ts}")
    n_input = parse_count(T, inputs["NINPUT"])
    n_max = parse_count(T, inputs["NMAX"])
    cellmin = parse_T(T, inputs["CELLMIN"])
    if mesh.geometry == "1D":
        BC = (str(inputs["LEFTBC"]).upper(), str(inputs["RIGHTBC"]).upper())
    else:
        BC = (str(inputs["LEFTBC"]).upper(), str(inputs["RIGHTBC"]).upper(), str(inputs["TOPBC"]).upper(), str(inputs["BOTTOMBC"]).upper())
    return SimVars(T(0), dt, dt0, k, dtmax, t_end, [], 0, n_input, n_max, cellmin, inputs["PAIRWISE"], BC, T, mesh.geometry)


@dataclass
class Simulation:
    """One deck bound to one engine; ``advance()`` runs one iteration of the reference's while loop."""
    inputs: dict
    mesh: MeshStruct
    simvars: SimVars
    particles: ParticleList
    rwvars: Optional[RWVars] = None
    log: List[dict] = field(default_factory=list)
    save_history: bool = True
    fused: bool = False

    @property
    def engine(self) -> _lib.Engine:
        return self.mesh.engine

    def fetch_history(self, clear: bool = True):
        """Append the snapshots the engine recorded (``engine.history_enable``) to the reference's lists
        ``mesh.temp_saved / matenergy_saved / radenergy_saved / energyincrease_saved`` (imc_tally.jl:58, :138-142), in
        the reference's array shapes, and leave the latest fields in ``mesh.temp`` etc. — one download at the end of a
        run (or every k steps) instead of three per step."""
        eng, mesh = self.engine, self.mesh
        n, _ = eng.history_count()
        if n == 0:
            return 0
        shape = (eng.cfg.nx, eng.cfg.ny) if eng.cfg.geometry == 2 else (eng.cfg.nx,)
        for name, lst in (("temp", mesh.temp_saved), ("matenergydens", mesh.matenergy_saved), ("radenergydens", mesh.radenergy_saved),
                          ("nrg_inc", mesh.energyincrease_saved)):
            h = eng.history(name).astype(np.float64)
            lst.extend(h[k].reshape(shape, order="F") for k in range(n))
        mesh.temp, mesh.matenergydens, mesh.radenergydens = mesh.temp_saved[-1], mesh.matenergy_saved[-1], mesh.radenergy_saved[-1]
        if clear:
            eng.history_clear()
        return n

    def done(self) -> bool:
        return not (self.simvars.t <= self.simvars.t_end)

    def advance(self) -> dict:
        inputs, mesh, sv, parts = self.inputs, self.mesh, self.simvars, self.particles
        rec = {"t": float(sv.t), "dt": float(sv.dt), "step": sv.step}
        if self.fused:  # imc_step: the whole stage sequence in one ABI call
            out = mesh.engine.step(float(sv.t), float(sv.dt), sv.n_input, float(sv.cellmin), sv.step)
            rec.update(out)
            sv.iterations = out["transport"]["segments_total"]
            mesh.totalenergy = out["source"]["totalenergy"]; mesh.totalenergydep = out["tally"]["totalenergydep"]
            if self.save_history:
                mesh.temp = mesh.engine.field("temp")
                mesh.matenergydens = mesh.engine.field("matenergydens")
                mesh.radenergydens = mesh.engine.field("radenergydens")
                mesh.temp_saved.append(mesh.temp.copy()); mesh.matenergy_saved.append(mesh.matenergydens.copy())
                mesh.radenergy_saved.append(mesh.radenergydens.copy())
        else:  # the reference's call sequence (MixedPrecisionIMC.jl:138-147 / :167-172)
            Update.update(inputs, mesh, sv)
            rec["source"] = Sourcing.sourcing(mesh, sv, parts)
            if mesh.geometry == "1D":
                if self.rwvars is not None:
                    rec["transport"] = Transport.MC_RW(mesh, sv, self.rwvars, parts)
                else:
                    rec["transport"] = Transport.MC(mesh, sv, parts)
            else:
                rec["transport"] = Transport.MC2D(mesh, sv, parts)
            Clean.clean(parts)
            rec["tally"] = Tally.tally(inputs, mesh, sv, parts, save=self.save_history)
            rec["energy"] = EnergyCheck.energychecker(inputs, mesh, sv, parts)
        timestep(str(inputs["TIMESTEPPING"]).upper(), sv)
        sv.step += 1
        self.log.append(rec)
        return rec


def setup(deck, library: Optional[_lib.ImcLib] = None, overrides: Optional[dict] = None, **cfg_overrides) -> Simulation:
    """readInputs + set_constants + mesh_generation + SimVars + engine (MixedPrecisionIMC.jl:78-134).

    ``deck`` is a deck file path or an already-parsed inputs dict; ``overrides`` replaces deck keywords
    (strings, as they would appear in the file) before the mesh is generated."""
    inputs = _deck.read_inputs(deck) if isinstance(deck, str) else dict(deck)
    if overrides:
        for k, v in overrides.items():
            inputs[k] = v
        if isinstance(inputs.get("PRECISION"), str):
            inputs["PRECISION"] = _deck._PRECISIONS[inputs["PRECISION"]]
    mesh = _deck.mesh_generation(inputs)
    simvars = make_simvars(inputs, mesh)
    eng = attach_engine(inputs, mesh, library, **cfg_overrides)
    rw = None
    if mesh.geometry == "1D" and str(inputs.get("RANDOMWALK", "FALSE")).upper() == "TRUE":
        rw = Transport.randomwalk_table(mesh)
    return Simulation(inputs, mesh, simvars, ParticleList(eng), rw)


def main(args, library: Optional[_lib.ImcLib] = None, max_steps: Optional[int] = None, quiet: bool = False) -> Optional[Simulation]:
    """``MixedPrecisionIMC.main(["deck.txt"])`` (MixedPrecisionIMC.jl:59-179) with the transport step on the GPU."""
    if not args:
        print("No input file provided, exiting... ")
        return None
    sim = setup(args[0], library)
    say = (lambda *a: None) if quiet else print
    say("Input file: ", args[0])
    n = 0
    while not sim.done() and (max_steps is None or n < max_steps):
        say("Time: ", sim.simvars.t)
        r = sim.advance()
        say("Total intial time-step energy ", r["source"]["emitted_sum"])
        say("The number of particles after sourcing is ", r["source"]["n_particles"])
        say("There were ", sim.simvars.iterations, " total iterations this time-step. ")
        say("Energy increase: ", r["tally"]["energy_increase"])
        say("Maximum mesh temperature is ", r["tally"]["max_temp"])
        say("Final total energy density ", r["tally"]["total_energy_density"])
        say("Total energy: ", r["source"]["totalenergy"], " Total energy deposition: ", r["tally"]["totalenergydep"],
            " Radiation energy change: ", r["energy"]["radenergy_change"], " Lost energy: ", r["energy"]["lostenergy"])
        say("The energy conservation error is: ", r["energy"]["energy_error"])
        n += 1
    return sim
