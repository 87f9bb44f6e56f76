set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/ab_kernel.py --cfg=synth4k --reps=9 ab/lib_head.so ab/lib_prev.so default > gpurun_out/r2n_ab_synth.txt 2>&1; cat gpurun_out/r2n_ab_synth.txt
timeout 900 python -m pytest tests/test_bvh_gpu.py tests/test_parity_gpu.py tests/test_random_scenes_gpu.py tests/test_edge_gpu.py tests/test_placed_gpu.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
python tools/one_frame.py synth4k 3 > gpurun_out/r2n_plain_synth4k.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r2n_prof_synth4k -f python tools/one_frame.py synth4k 3 > gpurun_out/r2n_ncu_synth4k.log 2>&1
