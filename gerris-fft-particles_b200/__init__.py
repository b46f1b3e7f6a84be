"""gerris-fft-particles_b200 -- B200-native Gerris Lagrangian particulate hot path.

The product is lib/libgfsb200.so (hand-written sm_100a CUDA kernels behind the
C-ABI of include/gfsb200.h) plus the per-dimension FttCell bridges
lib/libgfsb200_ftt{2D,3D}.so and the drop-in GModule source under host/.  The
Python modules here are the ctypes harness used by tests/ and bench.py:

  capi    ctypes binding of include/gfsb200.h (no compute, no fallback)
  worlds  synthetic trees / fields / particle clouds of the BASELINE configs
  multigpu  particle sharding + the two-way all-reduce (torch.distributed plumbing)

The directory name carries hyphens, so import it through
__graft_entry__.load_package(), which registers it as
`gerris_fft_particles_b200`.
"""
from . import capi  # noqa: F401
from . import worlds  # noqa: F401
from . import multigpu  # noqa: F401

__all__ = ["capi", "worlds", "multigpu"]
