set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2f_bench_n8.json 2> gpurun_out/r2f_bench_n8.err; echo "bench rc=$?"; tail -5 gpurun_out/r2f_bench_n8.err; head -c 2500 gpurun_out/r2f_bench_n8.json
