"""cgraytracing_b200 — B200-native (sm_100a) progressive photon mapping behind CGRayTracing's scene API.

The product is the CUDA library `libcgrt.so` (C ABI in include/cgrt.h); this package is the thin Python host
binding used by the tests and the benchmark, plus the scene presets. The C++ host mirror of the reference's
classes lives in cgraytracing_b200/host/.
"""
from .binding import CgrtError, Context, load_library  # noqa: F401
from .scene import PRESETS, RenderConfig, SceneDesc, preset  # noqa: F401
