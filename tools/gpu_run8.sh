set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for t in 0 16 48; do
RR_SUB_TAIL_16THS=$t timeout 600 $TR --master-port 2953$((t % 10)) bench.py --gpus 8 --steps 50 --warmup 5 --no-configs > gpurun_out/r2h_bench_n8_tail$t.json 2> gpurun_out/r2h_bench_n8_tail$t.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2h_bench_n8_tail$t.json'))
print('tail$t', 'step', d['ms_per_step'], 'kernel', d['kernel_ms'], 'same1', d['same_workload_1gpu_ms'], 'eff', d['efficiency_same_workload'], 'nvlink', d.get('nvlink_roofline'), 'e2e', d['e2e']['ms_per_frame'], d['frame_check'])
PY
done
timeout 900 $TR --master-port 29541 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2h_bench_n8.json 2> gpurun_out/r2h_bench_n8.err; echo "bench rc=$?"; tail -3 gpurun_out/r2h_bench_n8.err
