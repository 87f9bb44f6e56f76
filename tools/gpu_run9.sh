set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in trace8k synth4k; do
for s in 8 4 2 0; do for t in 0 16; do
RR_STATIC_16THS=$s RR_SUB_TAIL_16THS=$t timeout 300 python tools/shard_time.py $cfg 8 2>&1 | tail -1
done; done; done > gpurun_out/r2i_shard_time.txt 2>&1
cat gpurun_out/r2i_shard_time.txt
for s in 8 4 0; do RR_STATIC_16THS=$s RR_SUB_TAIL_16THS=0 timeout 300 python tools/shard_time.py trace4k 1 2>&1 | tail -1; done >> gpurun_out/r2i_shard_time.txt 2>&1
timeout 900 python tools/ab_kernel.py --cfg=march4k --reps=7 ab/lib_head.so default > gpurun_out/r2i_ab_march.txt 2>&1; cat gpurun_out/r2i_ab_march.txt
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_edge_gpu.py tests/test_random_scenes_gpu.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; tail -3 gpurun_out/r2i_pytest.log
