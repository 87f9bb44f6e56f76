set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/selftest_normalize.py ab/lib_rcp.so 28 > gpurun_out/r2x_selftest.txt 2>&1; echo "selftest rc=$?" >> gpurun_out/r2x_selftest.txt
python tools/selftest_normalize.py ab/lib_rcpni.so 26 >> gpurun_out/r2x_selftest.txt 2>&1; echo "selftest rc=$?" >> gpurun_out/r2x_selftest.txt
cat gpurun_out/r2x_selftest.txt
python tools/ab_kernel.py --cfg=trace4k,trace8k,synth4k,march4k --reps=25 ab/lib_noq.so ab/lib_rcp.so ab/lib_rcpni.so > gpurun_out/r2x_ab_rcp.txt 2>&1
cat gpurun_out/r2x_ab_rcp.txt
