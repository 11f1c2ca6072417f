"""shll_sve_cfd_b200 -- B200-native SHLL split-flux Euler time-march (drop-in for the hot path of
archembaud/shll-sve-cfd: Compute_F_from_P -> Update_U_from_F -> Compute_P_from_U).

The product is libshll_b200.so (hand-written sm_100a CUDA behind the C ABI in include/shll_b200.h) plus the C
host programs in host/.  This Python package is the thin host-side mirror used by tests and bench.py:

  capi      ctypes binding of the C ABI (fails loudly when the library or a GPU is missing -- no CPU fallback)
  programs  the reference programs' host-side logic (initial conditions, Compute_U_from_P, float clock,
            Save_Results), written with explicit float32/float64 steps so it is bit-identical to the C code
  slabs     one-process-per-GPU slab decomposition: torch.distributed is used only to exchange the CUDA-IPC
            descriptors and for barriers; the halo exchange itself is peer stores inside the step kernel
"""
from . import capi, programs  # noqa: F401

__all__ = ["capi", "programs"]
