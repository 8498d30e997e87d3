"""ramses_b200 -- B200 (sm_100a) per-level Godunov sweep for RAMSES (tatary/ramses).

Only what the hot path needs: `csrc/` (CUDA kernels + the C-ABI of include/ramses_gpu.h,
built in-tree into libramses_gpu.so) and `hydro`, a thin Python mirror of the reference's
module-global driver interface (amr_commons / hydro_commons + godunov_fine(ilevel) ...).
There is no CPU fallback: importing `ramses_b200.lib` fails loudly when the CUDA library
is missing, and every call fails when no GPU is usable.
"""
from .lib import load, build, RgpuError, Params, LevelInfo  # noqa: F401
from .hydro import AmrCommons, HydroGPU  # noqa: F401

__all__ = ["load", "build", "RgpuError", "Params", "LevelInfo", "AmrCommons", "HydroGPU"]
