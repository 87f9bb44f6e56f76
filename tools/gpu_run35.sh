cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r3n_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3n_pytest_gpu.log
tail -3 gpurun_out/r3n_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r3n_bench_n1.json 2> gpurun_out/r3n_bench_n1.err; echo "bench rc=$?"
cfg=march4k
python tools/one_frame.py $cfg 4 > gpurun_out/r3n_plain_$cfg.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -o gpurun_out/r3n_prof_$cfg -f python tools/one_frame.py $cfg 4 > gpurun_out/r3n_ncu_$cfg.log 2>&1
cat gpurun_out/r3n_plain_$cfg.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3n_bench_n1.json'))
print('value',round(d['value']),'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['ms_per_frame'])
for k,v in d.get('configs',{}).items(): print(k,'ms',round(v['ms_per_step'],4),'kernel',round(v['kernel_ms'],4),'frac',v['roofline'].get('frac'),'e2e',round(v['e2e']['ms_per_frame'],4),v['e2e']['check'])
PY
