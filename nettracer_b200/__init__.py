"""nettracer_b200 — B200-native intersect-and-shade path behind a C ABI (include/nettracer_b200.h).

Holds only what the hot path needs: csrc/ (CUDA kernels + the C-ABI library) and the host-side
scene/camera/renderer containers.  There is NO CPU fallback: importing `renderer` without the built
library, or rendering without an sm_100 GPU, raises."""
from . import abi  # noqa: F401
from .scene import Camera, Material, Scene, make_params  # noqa: F401

__all__ = ["abi", "Camera", "Material", "Scene", "make_params"]
