"""Annotated SASS listing of one kernel: address, dynamic warp-instruction count (ncu source csv), source line, instruction.
usage: python tools/sass_dump.py <obj> <function-substring> <ncu_source_sass.csv> [kernel_index] > listing.txt"""
import csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_lines import disasm, parse
per = parse(disasm(sys.argv[1]))
target = [f for f in per if sys.argv[2] in f][0]
ins = per[target]
rows = list(csv.reader(open(sys.argv[3])))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
k = int(sys.argv[4]) if len(sys.argv) > 4 else 0
h = rows[heads[k]]
end = heads[k + 1] - 1 if k + 1 < len(heads) else len(rows)
ci = h.index("Instructions Executed")
dyn = [int(r[ci]) for r in rows[heads[k] + 1:end] if len(r) > ci and r[ci].isdigit()]
assert len(dyn) == len(ins)
for (addr, s, li), n in zip(ins, dyn):
    print(f"{addr:06x} {n:9d}  {(li[0] + ':' + str(li[1])) if li else '?':24s} {s}")
