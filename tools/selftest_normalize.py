"""Device self-test of normalized() (shared-reciprocal divisions vs plain IEEE divisions): python tools/selftest_normalize.py [lib.so] [log2 n]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, sys.argv[1]) if len(sys.argv) > 1 else os.path.join(ROOT, "ray-rust_b200", "libray_rust_b200.so"))
lib.rr_last_error.restype = C.c_char_p
lib.rr_selftest_normalize.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 28)
for seed in (1, 0x9E3779B97F4A7C15, 20261018):
    bad = C.c_uint64(123)
    rc = lib.rr_selftest_normalize(0, n, seed, C.byref(bad))
    print(f"selftest_normalize n={n} seed={seed:#x}: rc={rc} mismatches={bad.value}", lib.rr_last_error().decode() if rc else "", flush=True)
    if rc or bad.value:
        sys.exit(1)
