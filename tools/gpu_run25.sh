cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
S0=$SECONDS
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r3o_bench_n$N.json 2> gpurun_out/r3o_bench_n$N.err; echo "bench rc=$? wall=$((SECONDS-S0))s"; tail -3 gpurun_out/r3o_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/r3o_bench_n$N.json'))
print('step', round(d['ms_per_step'],4), 'kernel', round(d['kernel_ms'],4), 'same1gpu', d.get('same_workload_1gpu_ms'), 'eff', d.get('efficiency_same_workload'), d.get('band_shares'), 'raw', d.get('nvlink_roofline'), d.get('frame_check'), d.get('per_rank_kernel_ms'), 'e2e', d['e2e'].get('ms_per_frame'), d['e2e'].get('check'))
print('alt', d.get('alt'))
for k,v in d.get('configs',{}).items(): print(k,'ms',v['ms_per_step'],'same1',v.get('same_workload_1gpu_ms'),'eff',v.get('efficiency_same_workload'), v.get('frame_check'))
PY
