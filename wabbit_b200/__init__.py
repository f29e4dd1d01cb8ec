"""wabbit_b200 -- B200-resident block hot path of WABBIT (ACM right-hand side, ghost synchronisation,
Runge-Kutta stages, wavelet decomposition/thresholding) behind a C ABI (include/wabbit_gpu.h).

Only what the path needs lives here:
  csrc/       hand-written sm_100a kernels + the C ABI (libwabbit_gpu.so) and the host forest tables (libwabbit_host.so)
  params.py   WABBIT .ini reader / parameter set
  forest.py   light data (hvy_active, levels, hvy_neighbor) from libwabbit_host.so
  solver.py   host-side mirror of the reference's tree-level routines
There is no CPU fallback: without libwabbit_gpu.so and a CUDA device, creating a solver raises.
"""
from .forest import Forest
from .params import IniFile, Params
from .solver import HVY_BLOCK, HVY_MASK, HVY_TMP, HVY_WORK, WabbitAbort, WabbitGPU

__all__ = ["Forest", "IniFile", "Params", "WabbitGPU", "WabbitAbort", "HVY_BLOCK", "HVY_WORK", "HVY_MASK", "HVY_TMP"]
