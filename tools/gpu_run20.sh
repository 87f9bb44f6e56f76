set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/ab_kernel.py --cfg=trace4k,trace8k --reps=25 ab/lib_base.so ab/lib_new.so ab/lib_noq.so ab/lib_init.so ab/lib_rev.so > gpurun_out/r2w_ab_trace.txt 2>&1
cat gpurun_out/r2w_ab_trace.txt
