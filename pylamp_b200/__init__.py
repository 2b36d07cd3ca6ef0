"""pylamp_b200 -- B200-native (sm_100a CUDA, fp64) implementation of PyLamp's per-timestep
hot path behind the reference's own module-level functions.  See DESIGN.md."""
__version__ = "0.1.0"
