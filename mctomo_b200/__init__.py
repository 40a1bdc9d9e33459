"""mctomo_b200 -- B200-native (sm_100a) implementation of MCTomo's surface-wave forward-modelling
hot path: Voronoi->grid nearest-nucleus assignment and per-column modal dispersion.

The product is the C-ABI shared library `libmctomo_b200.so` (include/mctomo_b200.h); `capi` binds
it for tests and benchmarks, `synth` generates the reference's own initial-model family.
"""
from . import capi, synth  # noqa: F401

__all__ = ["capi", "synth"]
