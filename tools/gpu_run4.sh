set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_gpu.py tests/test_placed_gpu.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; tail -5 gpurun_out/r2d_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; echo "bench rc=$?"; tail -5 gpurun_out/r2d_bench_n2.err; head -c 1500 gpurun_out/r2d_bench_n2.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2d_bench_n1.err; wc -l gpurun_out/r2d_bench_n1.json
