set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import ray_rust_b200 as rr; rr.ffi.load(); print('lib ok')" 2>&1 | tail -1
timeout 900 python tools/ab_kernel.py --cfg=trace4k,trace8k ab/lib_head.so default > gpurun_out/r2g_ab_trace.txt 2>&1; cat gpurun_out/r2g_ab_trace.txt
timeout 900 python tools/ab_kernel.py --cfg=synth4k --reps=9 ab/lib_head.so default ab/lib_t256.so ab/lib_t512.so ab/lib_sd12.so ab/lib_sd5.so ab/lib_leaf2.so > gpurun_out/r2g_ab_synth.txt 2>&1; cat gpurun_out/r2g_ab_synth.txt
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_bvh_gpu.py tests/test_random_scenes_gpu.py tests/test_edge_gpu.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -5 gpurun_out/r2g_pytest.log
