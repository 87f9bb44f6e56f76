set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.log
tail -5 gpurun_out/r2a_pytest_gpu.log
timeout 120 ab/ffma2_bench > gpurun_out/r2a_ffma2_bench.txt 2>&1; cat gpurun_out/r2a_ffma2_bench.txt
timeout 900 python tools/ab_kernel.py --cfg=trace4k,trace8k ab/lib_r1.so default ab/lib_scalar.so ab/lib_mb3.so > gpurun_out/r2a_ab_trace.txt 2>&1; cat gpurun_out/r2a_ab_trace.txt
timeout 900 python tools/ab_kernel.py --cfg=synth4k --reps=9 ab/lib_r1.so default ab/lib_sd0.so ab/lib_sd12.so ab/lib_mb3.so > gpurun_out/r2a_ab_synth.txt 2>&1; cat gpurun_out/r2a_ab_synth.txt
timeout 600 python tools/ab_kernel.py --cfg=march4k --reps=7 ab/lib_r1.so default > gpurun_out/r2a_ab_march.txt 2>&1; cat gpurun_out/r2a_ab_march.txt
python tools/one_frame.py trace4k 3 > gpurun_out/r2a_plain_trace.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r2a_prof_trace -f python tools/one_frame.py trace4k 3 > gpurun_out/r2a_ncu_trace.log 2>&1
python tools/one_frame.py synth4k 3 > gpurun_out/r2a_plain_synth.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r2a_prof_synth -f python tools/one_frame.py synth4k 3 > gpurun_out/r2a_ncu_synth.log 2>&1
ls -la gpurun_out | tail -15
