"""Loader for the package directory ``mixedprecisionimc.jl_b200`` (its name contains a dot, so a plain
``import`` statement cannot reach it).  ``import mpimc_b200`` gives the package; its submodules are then
importable as ``mpimc_b200.lib``, ``mpimc_b200.deck``, ``mpimc_b200.driver``, ``mpimc_b200.dist``."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "mixedprecisionimc.jl_b200")

_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
