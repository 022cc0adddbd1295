"""ndpp_b200 -- B200-native scattering-moment integrator behind NDPP's calc_scatt / calc_scattsab seam.

The compute path is the hand-written sm_100a CUDA library ndpp_b200/csrc/libndppgpu.so reached
through the C-ABI of include/ndppgpu.h.  There is no CPU fallback: every entry point raises if the
library is missing or no GPU is present.
"""
from . import ace, synth  # noqa: F401

__all__ = ["ace", "synth"]
