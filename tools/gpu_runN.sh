set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2p_bench_n$N.json 2> gpurun_out/r2p_bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/r2p_bench_n$N.err | cut -c1-300
