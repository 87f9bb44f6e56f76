set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/ab_kernel.py --cfg=synth4k --reps=9 ab/lib_r1.so ab/lib_sd0.so ab/lib_ns_sd0.so ab/lib_ns_sd8.so ab/lib_ns_sd16.so > gpurun_out/r2c_ab_synth.txt 2>&1; cat gpurun_out/r2c_ab_synth.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench rc=$?"; tail -5 gpurun_out/r2c_bench_n1.err; head -c 3000 gpurun_out/r2c_bench_n1.json
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; tail -5 gpurun_out/r2c_pytest.log
