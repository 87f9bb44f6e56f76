"""In-tree build of libray_rust_b200.so (CUDA kernels for sm_100a + the C ABI).

nvcc cross-compiles here without a GPU. -fmad=false is part of the numerics contract (bit parity
with Rust's unfused scalar f32, see csrc/rr_device.cuh); never add --use_fast_math.

Staleness is decided by CONTENT, not by modification time: the SHA-256 of every source under csrc/, of
include/rr_ffi.h and of the compiler flags is compiled into the library (rr_build_info(), include/rr_ffi.h) and
stored beside every object file. `build()` rebuilds whatever does not match, and `ffi.load()` refuses to hand out a
library whose embedded hash differs from the sources next to it (it rebuilds it first). Built files are kept out of
git history but travel to the GPU box; this is what guarantees that a shipped binary is the checked-in source.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libray_rust_b200.so")
SOURCES = ["rr_ffi.cu", "rr_trace.cu", "rr_march.cu", "rr_util.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _deps():
    out = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    out.append(os.path.join(HERE, "..", "include", "rr_ffi.h"))
    return out


def source_hash(extra_flags=()):
    """SHA-256 over the contents of every source the library is built from, and over the compiler flags."""
    h = hashlib.sha256()
    for d in _deps():
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as f:
            h.update(f.read())
        h.update(b"\0")
    h.update(" ".join(NVCC_FLAGS + list(extra_flags)).encode())
    return h.hexdigest()[:32]


def embedded_hash(lib_path=LIB):
    """The source hash compiled into a built library ('' when there is none)."""
    try:
        with open(lib_path, "rb") as f:
            blob = f.read()
    except OSError:
        return ""
    tag = b"rr_src_hash="
    i = blob.find(tag)
    return blob[i + len(tag):i + len(tag) + 32].decode("ascii", "replace") if i >= 0 else ""


def is_current(lib_path=LIB):
    return os.path.exists(lib_path) and embedded_hash(lib_path) == source_hash()


def build(force=False, verbose=False):
    """Build under an exclusive file lock: the ranks of a torchrun launch all call ffi.load() at the same time, and if
    the shipped library is stale they must not compile and link into the same files concurrently (one builds, the
    others wait and then find a current library)."""
    import fcntl

    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():
                return LIB
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose):
    want = source_hash()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        stamp = obj + ".hash"
        have = open(stamp).read().strip() if os.path.exists(stamp) else ""
        if not force and os.path.exists(obj) and have == want:
            return obj, "", False
        cmd = [_nvcc()] + NVCC_FLAGS + [f'-DRR_SRC_HASH="{want}"', "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(want)
        return obj, r.stderr, True

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _, _ in results]
    if verbose:
        for _, log, _ in results:
            sys.stderr.write(log)
    if force or any(c for _, _, c in results) or embedded_hash() != want:
        tmp = f"{LIB}.tmp.{os.getpid()}"
        cmd = [_nvcc(), "-shared", "-o", tmp] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB)  # atomic: a concurrent dlopen sees the old or the new library, never a partial one
    if embedded_hash() != want:
        raise RuntimeError("built library does not carry the hash of its sources")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
