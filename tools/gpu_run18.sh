set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29581 bench.py --gpus 8 --steps 50 --warmup 5 --no-configs --equal-shares > gpurun_out/r2t_bench_n8_equal.json 2> gpurun_out/r2t_bench_n8_equal.err; echo rc=$?
timeout 600 $TR --master-port 29582 bench.py --gpus 8 --steps 50 --warmup 5 --no-configs --shares 4,3 > gpurun_out/r2t_bench_n8_s43.json 2> gpurun_out/r2t_bench_n8_s43.err; echo rc=$?
timeout 900 $TR --master-port 29583 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2t_bench_n8.json 2> gpurun_out/r2t_bench_n8.err; echo rc=$?; tail -3 gpurun_out/r2t_bench_n8.err | cut -c1-300
python - <<'PY'
import json
for f in ('r2t_bench_n8_equal','r2t_bench_n8_s43','r2t_bench_n8'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, 'step', round(d['ms_per_step'],4), 'kernel', round(d['kernel_ms'],4), 'eff', round(d['efficiency_same_workload'],3), d.get('band_shares',{}).get('rank0_slots'), d.get('band_shares',{}).get('other_slots'), d.get('band_shares',{}).get('predicted_step_ms'), 'raw', round(d['nvlink_roofline']['raw_copy_ms'],4), d['frame_check'][:9], [round(x,4) for x in d['per_rank_kernel_ms']])
    except Exception as e: print(f,'ERR',e)
PY
