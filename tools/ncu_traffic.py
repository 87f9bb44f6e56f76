"""Refresh profiles/traffic.json (what bench.py reports as roofline.traffic / executed flops) from the round's ncu summaries:
    python tools/ncu_traffic.py r2e        (reads profiles/<tag>_{trace_default4k,march_default4k,trace_synthetic1024_4k}.json)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
names = {"default-4k-trace": "trace_default4k", "default-4k-march-glow": "march_default4k", "synthetic1024-4k-trace": "trace_synthetic1024_4k"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for wl, nm in names.items():
    rel = f"profiles/{tag}_{nm}.json"
    summ = json.load(open(os.path.join(ROOT, rel)))
    m = summ["metrics"]
    val = lambda k: float(m[k]["value"])
    rd = val("dram__bytes_read.sum") * UNIT[m["dram__bytes_read.sum"]["unit"]]
    wr = val("dram__bytes_write.sum") * UNIT[m["dram__bytes_write.sum"]["unit"]]
    cyc = val("sm__cycles_elapsed.max")
    per = lambda op: val(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")
    ffma2 = summ.get("thread_inst_fp32", {}).get("FFMA2", 0)  # not in ncu's op_ffma counter; two FP32 operations per thread
    ex = (per("fadd") + per("fmul") + 2 * per("ffma")) * cyc + 2 * ffma2
    out[wl] = {
        "dram_bytes_per_launch": int(rd + wr),
        "note": f"ncu --set full ({rel[:-5]}.md): dram__bytes_read {rd/1e6:.2f} MB + dram__bytes_write {wr/1e6:.2f} MB per launch "
                f"(kernel {m['gpu__time_duration.sum']['value']} {m['gpu__time_duration.sum']['unit']} under ncu); the RGB8 frame is stored once and is "
                "still (partly) resident in the 126 MB L2 when the kernel ends",
        "executed_flops_per_launch": int(ex),
        "executed_note": f"ncu ({rel}): thread-level (FADD + FMUL + 2 x FFMA) per elapsed cycle x sm__cycles_elapsed.max + 2 x thread-level FFMA2 of the SASS page "
                         f"({ffma2} packed instructions: two FP32 operations each, not in ncu's op_ffma counter)",
    }
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
