"""ray-rust_b200 — B200-native per-pixel tracing path of ray-rust behind a C ABI.

Import name: `ray_rust_b200` (the directory is `ray-rust_b200/`; the root-level shim
`ray_rust_b200.py` maps one onto the other). Contents:
  csrc/   CUDA kernels for sm_100a + the C ABI implementation (include/rr_ffi.h)
  host/   C++ host layer mirroring render.rs / main.rs (RenderEnv, YAML, CLI)
  ffi.py, scene.py   Python binding of the C ABI and scene model mirror (tests, bench)
"""
from . import ffi  # noqa: F401
from .scene import (  # noqa: F401
    MAX_REFLECTIONS, MAX_REFRACTIONS, Camera, DeserializeError, DeviceScene, Quat, RenderColor, RenderEnv,
    RenderFloor, RenderMaterial, RenderSphere, SplitMix64, default_scene, frame_rows, render, synthetic_scene,
)
