"""ctypes mirror of include/rr_ffi.h and loader of the CUDA library.

This is one of the two bindings of the C ABI that are built and tested in this repo (the other is
the C++ host layer in host/). There is no CPU fallback: if libray_rust_b200.so is missing the
loader raises, it never routes to another implementation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RAY_RUST_B200_LIB selects another build of the same library (A/B runs of kernel variants)
LIB_PATH = os.environ.get("RAY_RUST_B200_LIB") or os.path.join(_HERE, "libray_rust_b200.so")

RR_OK = 0
RR_ERR_BAD_ARG = -1
RR_ERR_CUDA = -2
RR_ERR_OOM = -3
RR_ERR_UNSUPPORTED = -4

RR_SPHERE, RR_FLOOR = 0, 1
RR_UV_XY, RR_UV_YZ, RR_UV_ZX, RR_UV_LL = 0, 1, 2, 3
RR_SOLID, RR_CHECKERBOARD, RR_REPEATED_GRADATION = 0, 1, 2
RR_NEAREST, RR_BILINEAR = 0, 1
RR_BG_BGCOLOR, RR_BG_BLACK = 0, 1


class rr_material(C.Structure):
    _fields_ = [
        ("diffuse", C.c_float * 3),
        ("specular", C.c_float * 3),
        ("pn", C.c_int32),
        ("t", C.c_float),
        ("n", C.c_float),
        ("glow_dist", C.c_float),
        ("frac", C.c_float * 3),
        ("pattern", C.c_int32),
        ("pattern_scale", C.c_float),
        ("pattern_angle_scale", C.c_float),
        ("texture", C.c_int32),
        ("texture_filter", C.c_int32),
    ]


class rr_object(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("material", C.c_int32),
        ("uvmap", C.c_int32),
        ("r", C.c_float),
        ("org", C.c_float * 3),
        ("face_normal", C.c_float * 3),
    ]


class rr_texture(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgb8", C.POINTER(C.c_uint8))]


class rr_scene_desc(C.Structure):
    _fields_ = [
        ("n_objects", C.c_uint32),
        ("objects", C.POINTER(rr_object)),
        ("n_materials", C.c_uint32),
        ("materials", C.POINTER(rr_material)),
        ("n_textures", C.c_uint32),
        ("textures", C.POINTER(rr_texture)),
    ]


class rr_frame_params(C.Structure):
    _fields_ = [
        ("xres", C.c_int32),
        ("yres", C.c_int32),
        ("xfov", C.c_float),
        ("yfov", C.c_float),
        ("cam_position", C.c_float * 3),
        ("cam_rotation", C.c_float * 4),
        ("light", C.c_float * 3),
        ("use_raymarching", C.c_int32),
        ("glow_enabled", C.c_int32),
        ("glow_effect", C.c_float),
        ("max_reflections", C.c_int32),
        ("max_refractions", C.c_int32),
        ("bg_kind", C.c_int32),
        ("band_rows", C.c_int32),
        ("band_index", C.c_int32),
        ("band_count", C.c_int32),
        ("band_span", C.c_int32),
    ]


class rr_ray_counts(C.Structure):
    _fields_ = [
        ("pixels", C.c_uint64),
        ("primary", C.c_uint64),
        ("reflect", C.c_uint64),
        ("refract", C.c_uint64),
        ("shadow", C.c_uint64),
        ("object_tests", C.c_uint64),
        ("march_steps", C.c_uint64),
        ("bg_evals", C.c_uint64),
        ("sphere_tests", C.c_uint64),
        ("sphere_hits", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}

    def rays(self):
        """one ray = one scene-level raycast() / raymarch_single() call."""
        return int(self.primary + self.reflect + self.refract + self.shadow)


# every symbol include/rr_ffi.h declares: name -> (restype, argtypes)
_P = C.c_void_p
PROTOTYPES = {
    "rr_abi_version": (C.c_int, []),
    "rr_last_error": (C.c_char_p, []),
    "rr_build_info": (C.c_char_p, []),
    "rr_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rr_scene_create": (C.c_int, [C.POINTER(rr_scene_desc), C.c_int, C.POINTER(_P)]),
    "rr_scene_destroy": (C.c_int, [_P]),
    "rr_scene_set_culling": (C.c_int, [_P, C.c_int]),
    "rr_frame_rows": (C.c_int, [C.POINTER(rr_frame_params), C.POINTER(C.c_int32)]),
    "rr_render_rgb8": (C.c_int, [_P, C.POINTER(rr_frame_params), _P, C.c_size_t]),
    "rr_render_rgb8_async": (C.c_int, [_P, C.POINTER(rr_frame_params), _P, C.c_size_t, C.POINTER(C.c_int32)]),
    "rr_render_wait": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_float)]),
    "rr_render_f32": (C.c_int, [_P, C.POINTER(rr_frame_params), _P]),
    "rr_render_rgb8_device": (C.c_int, [_P, C.POINTER(rr_frame_params), _P, C.c_size_t, _P]),
    "rr_render_f32_device": (C.c_int, [_P, C.POINTER(rr_frame_params), _P, _P]),
    "rr_render_count": (C.c_int, [_P, C.POINTER(rr_frame_params), _P, C.c_size_t, C.POINTER(rr_ray_counts)]),
    "rr_bands_unpack_device": (C.c_int, [C.POINTER(rr_frame_params), _P, C.c_size_t, _P, _P]),
    "rr_render_rgb8_placed_device": (C.c_int, [_P, C.POINTER(rr_frame_params), _P, C.c_size_t, _P]),
    "rr_render_rgb8_placed": (C.c_int, [_P, C.POINTER(rr_frame_params), _P, C.c_size_t]),
    "rr_render_rgb8_placed_signal_device": (C.c_int, [_P, C.POINTER(rr_frame_params), _P, C.c_size_t, _P, C.c_uint32, _P]),
    "rr_fence_wait_device": (C.c_int, [C.c_int, _P, C.c_int32, C.c_uint32, C.c_uint32, _P, _P]),
    "rr_fence_signal_device": (C.c_int, [C.c_int, _P, C.c_uint32, _P]),
    "rr_device_memset": (C.c_int, [C.c_int, _P, C.c_int, C.c_size_t]),
    "rr_device_read": (C.c_int, [C.c_int, _P, _P, C.c_size_t]),
    "rr_device_copy": (C.c_int, [C.c_int, _P, _P, C.c_size_t, _P]),
    "rr_device_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(_P)]),
    "rr_device_free": (C.c_int, [C.c_int, _P]),
    "rr_ipc_export": (C.c_int, [_P, _P]),
    "rr_ipc_open": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "rr_ipc_close": (C.c_int, [C.c_int, _P]),
    "rr_host_register": (C.c_int, [_P, C.c_size_t]),
    "rr_host_unregister": (C.c_int, [_P]),
    "rr_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "rr_host_free": (C.c_int, [_P]),
    "rr_last_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "rr_fp32_peak_tflops": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "rr_selftest_normalize": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
}

_lib = None


class RrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rr error {code}: {msg}")
        self.code = code


def load():
    """Load libray_rust_b200.so (built in-tree by build.py). Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if LIB_PATH == os.path.join(_HERE, "libray_rust_b200.so"):
        # The in-tree binary must be the checked-in source: built files are not in git history but do travel to the
        # GPU box, so a stale one could otherwise be loaded silently. Compare the hash compiled into it with the
        # sources next to it and rebuild when they differ (nvcc, a few minutes, once).
        import importlib.util
        import sys

        spec = importlib.util.spec_from_file_location("rr_build", os.path.join(_HERE, "build.py"))
        b = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(b)
        if not b.is_current(LIB_PATH):
            sys.stderr.write("ray_rust_b200: libray_rust_b200.so is missing or does not match csrc/ (content hash); rebuilding\n")
            b.build()
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'`. There is no CPU fallback in this package."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != RR_OK:
        msg = load().rr_last_error()
        raise RrError(code, msg.decode("utf-8", "replace") if msg else "")
