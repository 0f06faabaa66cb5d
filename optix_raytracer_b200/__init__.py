"""optix_raytracer_b200 — B200-native (sm_100a) replacement for the OptiX launch behind the
optixPathTracer / optixMultiGPU / optixRaycasting samples of awegsche/OptiX_Raytracer.

The product is libb200rt.so (hand-written CUDA behind the C ABI of include/b200rt.h); this package
holds its sources (csrc/), the ctypes binding (_lib) and a host-side mirror of the samples' launch
plumbing (host).  Importing the package does not need a GPU; using it does, and there is no CPU path.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1"
