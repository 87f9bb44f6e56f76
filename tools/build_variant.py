"""Build a variant of libray_rust_b200.so with extra nvcc flags for A/B runs:
    python tools/build_variant.py gpurun_out/lib_unordered.so -DRR_BVH_ORDERED=0
Use it with RAY_RUST_B200_LIB=<path> (ray-rust_b200/ffi.py) or tools/ab_lib.py."""
import importlib.util, os, subprocess, sys, tempfile
from concurrent.futures import ThreadPoolExecutor
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("rr_build", os.path.join(root, "ray-rust_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
out, flags = sys.argv[1], sys.argv[2:]
tmp = tempfile.mkdtemp()
def cc(src):
    obj = os.path.join(tmp, src.replace(".cu", ".o"))
    subprocess.check_call([b._nvcc()] + [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + flags + ["-c", os.path.join(b.CSRC, src), "-o", obj])
    return obj
with ThreadPoolExecutor(4) as ex:
    objs = list(ex.map(cc, b.SOURCES))
os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
subprocess.check_call([b._nvcc(), "-shared", "-o", out] + objs)
print(out)
