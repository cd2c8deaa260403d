"""subsweep_b200 -- B200-native (sm_100a) drop-in for subsweep's sweep + chemistry hot path.

The product is ``lib/libsubsweep_b200.so`` (hand-written CUDA behind the C ABI of
``include/subsweep_b200.h``).  The Python modules only marshal arrays to it:

* :mod:`subsweep_b200.capi`   ctypes binding (twin of the Rust FFI crate in INTEGRATION.md)
* :mod:`subsweep_b200.sweep`  ``SweepParameters`` / ``Directions`` / ``Sweep`` / ``SweepPlugin``
* :mod:`subsweep_b200.grid`   flat-grid producers (Cartesian, Voronoi via Qhull) -- host preprocessing
* :mod:`subsweep_b200.snapshot`  snapshot output of the per-particle components (names, units, layout of the reference)
* :mod:`subsweep_b200.distributed`  NCCL hooks for direction sharding (all-reduce, reduce-scatter / all-gather)
* :mod:`subsweep_b200.build`  nvcc build of the library
"""
from .grid import FlatGrid  # noqa: F401
from .sweep import Directions, Sweep, SweepParameters, SweepPlugin, direction_shard  # noqa: F401

__all__ = ["FlatGrid", "Directions", "Sweep", "SweepParameters", "SweepPlugin", "direction_shard"]
