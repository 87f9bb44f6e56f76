set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python tools/ab_kernel.py --cfg=march4k --reps=7 ab/lib_head.so default ab/lib_conv2.so ab/lib_conv2mb7.so > gpurun_out/r2l_ab_march.txt 2>&1; cat gpurun_out/r2l_ab_march.txt
