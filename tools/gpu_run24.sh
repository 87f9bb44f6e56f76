cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S0=$SECONDS; timeout 900 python bench.py > gpurun_out/r3a_bench_n1.json 2> gpurun_out/r3a_bench_n1.err; echo "bench rc=$? wall=$((SECONDS-S0))s"; tail -3 gpurun_out/r3a_bench_n1.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3a_bench_n1.json'))
print('value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['kernel_ms'],'frac',d['roofline']['frac'],'e2e',d['e2e']['ms_per_frame'],d['e2e']['check'],'clocks',d['clocks'])
for k,v in d.get('configs',{}).items(): print(k,'ms',v['ms_per_step'],'kernel',v['kernel_ms'],'frac',v['roofline'].get('frac'),v['roofline'].get('frac_executed'),'e2e',v['e2e']['ms_per_frame'],v['e2e']['check'])
PY
