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
    m = json.load(open(os.path.join(ROOT, rel)))["metrics"]
    val = lambda k: float(m[k]["value"])
    rd = val("dram__bytes_read.sum") * UNIT[m["dram__bytes_read.sum"]["unit"]]
    wr = val("dram__bytes_write.sum") * UNIT[m["dram__bytes_write.sum"]["unit"]]
    cyc = val("sm__cycles_elapsed.max")
    per = lambda op: val(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")
    ex = (per("fadd") + per("fmul") + 2 * per("ffma")) * cyc
    out[wl] = {
        "dram_bytes_per_launch": int(rd + wr),
        "note": f"ncu --set full ({rel[:-5]}.md): dram__bytes_read {rd/1e6:.2f} MB + dram__bytes_write {wr/1e6:.2f} MB per launch "
                f"(kernel {m['gpu__time_duration.sum']['value']} {m['gpu__time_duration.sum']['unit']} under ncu); the RGB8 frame is stored once and is "
                "still (partly) resident in the 126 MB L2 when the kernel ends",
        "executed_flops_per_launch": int(ex),
        "executed_note": f"ncu ({rel}): thread-level (FADD + FMUL + 2 x FFMA/FFMA2-slot) per elapsed cycle x sm__cycles_elapsed.max",
    }
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
