"""sdetools_b200 — B200-native ensemble SDE integrator behind the SDEtools sample-path API.

Host-side mirror of the reference interface (api.py) over the C ABI of
libsdegpu.so (include/sdegpu.h).  No CPU fallback: without the built library
and an sm_100 GPU every compute call raises.
"""
from . import catalogue
from .api import (CrossVariation, QuadraticVariation, dCIR, dOU, pCIR, pOU, euler, heun, itointegral, moments_from_power_sums, rBM,
                  rBrownianBridge, rCIR, rOU, rvBM, rvBM_functionals, stochint)
from .engine import Engine, default_engine
from .laws import dLinSDE, linear_exact, lyap

__all__ = ["catalogue", "euler", "heun", "stochint", "itointegral", "QuadraticVariation", "CrossVariation", "rBM",
           "rvBM", "rvBM_functionals", "rBrownianBridge", "rOU", "rCIR", "dOU", "pOU", "dCIR", "pCIR", "dLinSDE", "lyap", "linear_exact", "Engine", "default_engine", "moments_from_power_sums"]
