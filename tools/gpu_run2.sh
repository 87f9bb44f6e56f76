set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/ab_kernel.py --cfg=trace4k,trace8k ab/lib_r1.so default ab/lib_plain.so ab/lib_plain_ng.so > gpurun_out/r2b_ab_trace.txt 2>&1; cat gpurun_out/r2b_ab_trace.txt
timeout 900 python tools/ab_kernel.py --cfg=synth4k --reps=9 ab/lib_r1.so default ab/lib_sd0.so ab/lib_sd4.so > gpurun_out/r2b_ab_synth.txt 2>&1; cat gpurun_out/r2b_ab_synth.txt
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_bvh_gpu.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
