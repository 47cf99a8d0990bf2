"""mixedprecisionimc.jl_b200 — B200-native Implicit Monte Carlo transport-step engine.

Drop-in for the transport step of simonbutson/MixedPrecisionIMC.jl (update -> source -> track ->
clean -> tally), behind the C ABI declared in include/imc.h and implemented as hand-written sm_100a
CUDA in csrc/.  This Python package is only the host-side mirror used where Julia is unavailable:

  lib     ctypes binding of include/imc.h (loads libimc_b200.so; no CPU fallback)
  deck    look-alike of the reference's deck parser / mesh generator (host side, stays Julia in production)
  driver  mirror of MixedPrecisionIMC.main and of the per-stage call sites (Update.update, ...)
  dist    one-process-per-GPU particle sharding over torch.distributed

The directory name contains a dot, so import it through the root-level loader module ``mpimc_b200``.
"""
from . import lib, deck, decks, driver  # noqa: F401
from .lib import Config, Engine, ImcLib, ImcError, cuda_lib  # noqa: F401
from .driver import main, setup, Simulation  # noqa: F401

__version__ = "0.1.0"
