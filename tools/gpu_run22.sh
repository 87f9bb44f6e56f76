set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/selftest_normalize.py ab/lib_uv.so 28 > gpurun_out/r2y_selftest.txt 2>&1; echo "selftest rc=$?" >> gpurun_out/r2y_selftest.txt
cat gpurun_out/r2y_selftest.txt
python tools/ab_kernel.py --cfg=trace4k,trace8k --reps=25 ab/lib_rcp.so ab/lib_uv.so ab/lib_mb3.so ab/lib_mb5.so > gpurun_out/r2y_ab.txt 2>&1
python tools/ab_kernel.py --cfg=march4k --reps=15 ab/lib_noq.so ab/lib_uv.so ab/lib_uvmp.so >> gpurun_out/r2y_ab.txt 2>&1
cat gpurun_out/r2y_ab.txt
