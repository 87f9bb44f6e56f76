set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_placed_gpu.py -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; tail -3 gpurun_out/r2s_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 3 --no-configs --shares 3,2 > gpurun_out/r2s_bench_n2_shares.json 2> gpurun_out/r2s_bench_n2_shares.err; echo rc=$?; tail -3 gpurun_out/r2s_bench_n2_shares.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 20 --warmup 3 --no-configs > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err; echo rc=$?; tail -3 gpurun_out/r2s_bench_n2.err | cut -c1-300
