"""A/B several builds of the library on the same box: python tools/ab_lib.py lib1.so lib2.so ... [--cfg=trace4k,synth4k] (runs quick_time in subprocesses, interleaved)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = [a for a in sys.argv[1:] if not a.startswith("--cfg=")]
cfgs = [a[6:].split(",") for a in sys.argv[1:] if a.startswith("--cfg=")]
cfgs = cfgs[0] if cfgs else ["trace4k", "march4k"]
for rep in range(2):
    for lib in libs:
        env = dict(os.environ)
        if lib != "default":
            env["RAY_RUST_B200_LIB"] = os.path.join(root, lib)
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "quick_time.py"), *cfgs], env=env, capture_output=True, text=True).stdout
        for l in out.splitlines():
            if "kernel ms" in l:
                print(lib, l[:120], flush=True)
