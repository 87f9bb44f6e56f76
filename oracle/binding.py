"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs. The product package (ray-rust_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ray_rust_b200 import ffi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_F3 = C.POINTER(C.c_float)


def build(force=False):
    src = os.path.join(_HERE, "rr_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB_PATH)
    lib.oracle_render.restype = C.c_int
    lib.oracle_render.argtypes = [C.POINTER(ffi.rr_scene_desc), C.POINTER(ffi.rr_frame_params), C.c_int,
                                  C.c_void_p, C.c_void_p, C.POINTER(ffi.rr_ray_counts), C.c_void_p]
    lib.oracle_fmod.restype = C.c_float
    lib.oracle_fmod.argtypes = [C.c_float, C.c_float]
    lib.oracle_imod.restype = C.c_int32
    lib.oracle_imod.argtypes = [C.c_int32, C.c_int32]
    lib.oracle_umod.restype = C.c_uint32
    lib.oracle_umod.argtypes = [C.c_uint32, C.c_uint32]
    lib.oracle_fimod.restype = None
    lib.oracle_fimod.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
    lib.oracle_powi.restype = C.c_float
    lib.oracle_powi.argtypes = [C.c_float, C.c_int32]
    lib.oracle_quantize.restype = C.c_uint8
    lib.oracle_quantize.argtypes = [C.c_float]
    lib.oracle_sphere_raycast.restype = C.c_float
    lib.oracle_sphere_raycast.argtypes = [_F3, C.c_float, _F3, _F3, C.c_float, C.c_uint32]
    lib.oracle_floor_raycast.restype = C.c_float
    lib.oracle_floor_raycast.argtypes = [_F3, _F3, _F3, _F3, C.c_float]
    lib.oracle_sphere_distance.restype = C.c_float
    lib.oracle_sphere_distance.argtypes = [_F3, C.c_float, _F3]
    lib.oracle_floor_distance.restype = C.c_float
    lib.oracle_floor_distance.argtypes = [_F3, _F3, _F3]
    for name in ("oracle_scale_pixel", "oracle_add_pixel", "oracle_quat_from_pyr", "oracle_quat_mul",
                 "oracle_quat_transform", "oracle_quat_slerp", "oracle_normalize", "oracle_primary_ray",
                 "oracle_bgcolor", "oracle_get_uv", "oracle_lookup_texture", "oracle_trace_pixel"):
        getattr(lib, name).restype = None
    lib.oracle_scale_pixel.argtypes = [C.c_float, C.POINTER(C.c_uint8), _F3]
    lib.oracle_add_pixel.argtypes = [_F3, _F3, _F3]
    lib.oracle_quat_from_pyr.argtypes = [_F3, _F3]
    lib.oracle_quat_mul.argtypes = [_F3, _F3, _F3]
    lib.oracle_quat_transform.argtypes = [_F3, _F3, _F3]
    lib.oracle_quat_slerp.argtypes = [_F3, _F3, C.c_float, _F3]
    lib.oracle_normalize.argtypes = [_F3, _F3]
    lib.oracle_primary_ray.argtypes = [C.POINTER(ffi.rr_frame_params), C.c_int32, C.c_int32, _F3]
    lib.oracle_bgcolor.argtypes = [_F3, _F3, _F3]
    lib.oracle_get_uv.argtypes = [C.POINTER(ffi.rr_material), _F3, C.c_int32, _F3]
    lib.oracle_lookup_texture.argtypes = [C.POINTER(ffi.rr_material), C.POINTER(ffi.rr_texture), C.c_float,
                                          C.c_float, _F3]
    lib.oracle_trace_pixel.argtypes = [C.POINTER(ffi.rr_scene_desc), C.POINTER(ffi.rr_frame_params), C.c_int32,
                                       C.c_int32, _F3]
    _lib = lib
    return lib


def fa(*vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def render(ren, params=None, threads=1, want_f32=False, want_u8=True, want_counts=False, want_tags=False):
    """Render `ren` (a ray_rust_b200.RenderEnv) with the oracle. Returns a dict."""
    from ray_rust_b200 import frame_rows

    lib = load()
    flat = ren.flatten()
    if params is None:
        params = ren.frame_params()
    rows = frame_rows(params)
    out = {}
    f32buf = np.empty((rows, params.xres, 3), dtype=np.float32) if want_f32 else None
    u8buf = np.empty((rows, params.xres, 3), dtype=np.uint8) if want_u8 else None
    tags = np.zeros((rows, params.xres), dtype=np.uint32) if want_tags else None
    counts = ffi.rr_ray_counts() if (want_counts or want_tags) else None
    rc = lib.oracle_render(C.byref(flat.desc), C.byref(params), int(threads),
                           f32buf.ctypes.data_as(C.c_void_p) if f32buf is not None else None,
                           u8buf.ctypes.data_as(C.c_void_p) if u8buf is not None else None,
                           C.byref(counts) if counts is not None else None,
                           tags.ctypes.data_as(C.c_void_p) if tags is not None else None)
    if rc != 0:
        raise RuntimeError(f"oracle_render failed: {rc}")
    out["f32"], out["u8"], out["tags"], out["counts"] = f32buf, u8buf, tags, counts
    return out
