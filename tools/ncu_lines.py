"""Per-CUDA-source-line totals of an ncu report taken with --import-source on (no object file needed):
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
Prints warp instructions executed, average active threads and stall samples per source line, ranked."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ci = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples")
    elif hdr and r[0].isdigit():
        try:
            out.append((int(r[ci]), int(r[ti]), int(r[si]), cur_file, int(r[0]), r[1].strip()[:90]))
        except ValueError:
            pass
tot = sum(o[0] for o in out); tots = sum(o[2] for o in out)
print(f"total warp instructions {tot}, samples {tots}")
for n, t, s, f, l, src in sorted(out, reverse=True)[:top]:
    print(f"{n:11d} {100*n/tot:5.1f}%  lanes {t/max(n,1):5.1f}  samples {100*s/max(tots,1):5.1f}%  {f}:{l}  {src}")
