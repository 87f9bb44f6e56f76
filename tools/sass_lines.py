"""Static (and, with an ncu source csv, dynamic) SASS instruction counts per CUDA source line.

usage: python tools/sass_lines.py <obj-or-cubin> <function-substring> [ncu_source.csv kernel_index]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def disasm(path):
    if not path.endswith(".cubin"):
        d = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=d, capture_output=True)
        cubins = [os.path.join(d, f) for f in os.listdir(d) if "sm_100" in f]
        out = ""
        for c in cubins:
            out += subprocess.run(["nvdisasm", "-g", "-c", c], capture_output=True, text=True).stdout
        return out
    return subprocess.run(["nvdisasm", "-g", "-c", path], capture_output=True, text=True).stdout


def parse(txt):
    fn, cur, per = None, None, collections.defaultdict(list)
    for l in txt.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn, cur = m.group(1), None
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and fn:
            per[fn].append((int(m.group(1), 16), m.group(2), cur))
    return per


def main():
    per = parse(disasm(sys.argv[1]))
    target = [f for f in per if sys.argv[2] in f][0]
    ins = per[target]
    dyn = None
    if len(sys.argv) > 3:
        rows = list(csv.reader(open(sys.argv[3])))
        heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
        k = int(sys.argv[4]) if len(sys.argv) > 4 else 0
        h = rows[heads[k]]
        end = heads[k + 1] - 1 if k + 1 < len(heads) else len(rows)
        ci = h.index("Instructions Executed")
        dyn = [int(r[ci]) for r in rows[heads[k] + 1:end] if len(r) > ci and r[ci].isdigit()]
        assert len(dyn) == len(ins), (len(dyn), len(ins))
    stat, dynagg = collections.Counter(), collections.Counter()
    for i, (addr, s, li) in enumerate(ins):
        key = (li[0], li[1]) if li else ("?", 0)
        stat[key] += 1
        if dyn:
            dynagg[key] += dyn[i]
    src = {}
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ray-rust_b200", "csrc")
    for f in os.listdir(here):
        src[f] = open(os.path.join(here, f)).read().split("\n")
    print(target, "static SASS:", len(ins), "dynamic warp-inst:", sum(dyn) if dyn else "-")
    tot = sum(dyn) if dyn else len(ins)
    rank = dynagg if dyn else stat
    for (f, l), n in rank.most_common(60):
        line = src[f][l - 1].strip()[:80] if f in src and l - 1 < len(src[f]) else ""
        print(f"{n:11d} {100 * n / tot:5.1f}%  static {stat[(f, l)]:4d}  {f}:{l}  {line}")


if __name__ == "__main__":
    main()
