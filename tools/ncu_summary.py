"""Summarise an .ncu-rep (ncu --set full --import-source on) into profiles/<name>.md + .json.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_trace_default4k [kernel_index]
Reads the report here (no GPU needed): `ncu -i rep --page raw --csv` and `--page source --csv`.
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__sass_branch_targets_threads_divergent.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.max",
    "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
]
STALLS = "smsp__average_warps_issue_stalled_"


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    raw = [r for r in raw if r and not r[0].startswith("==")]
    hdr, units, rows = raw[0], raw[1], raw[2:]
    row = rows[min(kidx, len(rows) - 1)]
    col = {h: i for i, h in enumerate(hdr)}
    summary = {"kernel": row[col["Kernel Name"]], "metrics": {}, "stalls_per_issue": {}}
    for m in METRICS:
        if m in col:
            summary["metrics"][m] = {"value": row[col[m]], "unit": units[col[m]]}
    for h in hdr:
        if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"):
            summary["stalls_per_issue"][h[len(STALLS):-len("_per_issue_active.ratio")]] = float(row[col[h]])
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    heads = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    ops = collections.Counter()
    total = 0
    n_sass = 0
    if heads:
        k = min(kidx, len(heads) - 1)
        h = src[heads[k]]
        end = heads[k + 1] - 1 if k + 1 < len(heads) else len(src)
        ci, si = h.index("Instructions Executed"), h.index("Source")
        ti = h.index("Predicated-On Thread Instructions Executed") if "Predicated-On Thread Instructions Executed" in h else None
        tops = collections.Counter()
        for r in src[heads[k] + 1:end]:
            if len(r) <= ci:
                continue
            try:
                n = int(r[ci])
            except ValueError:
                continue
            n_sass += 1
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
            ops[m.group(2).split(".")[0] if m else "?"] += n
            total += n
            if ti is not None and len(r) > ti and r[ti].isdigit():
                tops[m.group(2).split(".")[0] if m else "?"] += int(r[ti])
    summary["sass_instructions"] = n_sass
    summary["warp_instructions_executed"] = total
    summary["opcode_mix_pct"] = {o: round(100.0 * n / total, 2) for o, n in ops.most_common(25)} if total else {}
    if heads:
        # thread-level (predicated-on) counts of the FP32 opcodes: ncu's op_ffma counter does not include the packed FFMA2
        summary["thread_inst_fp32"] = {o: tops.get(o, 0) for o in ("FADD", "FMUL", "FFMA", "FFMA2", "FMNMX", "FMNMX3", "MUFU")}
    json.dump(summary, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary: {summary['kernel']}\n\nsource: `{rep}` (ncu --set full --clock-control none --import-source on)\n\n")
        f.write("| metric | value | unit |\n|---|---|---|\n")
        for m, v in summary["metrics"].items():
            f.write(f"| {m} | {v['value']} | {v['unit']} |\n")
        f.write("\n## warp stall reasons (warps stalled per issue-active cycle)\n\n| reason | ratio |\n|---|---|\n")
        for k2, v in sorted(summary["stalls_per_issue"].items(), key=lambda kv: -kv[1]):
            f.write(f"| {k2} | {v:.3f} |\n")
        f.write(f"\n## SASS\n\n{n_sass} SASS instructions in the kernel; {total} warp-instructions executed.\n\n| opcode | % of executed |\n|---|---|\n")
        for o, pct in summary["opcode_mix_pct"].items():
            f.write(f"| {o} | {pct} |\n")
    print("wrote", out + ".md")


if __name__ == "__main__":
    main()
