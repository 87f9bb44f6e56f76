set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/ab_kernel.py --cfg=synth4k --reps=9 default ab/lib_leaf3.so ab/lib_leaf6.so ab/lib_leaf7.so > gpurun_out/r2o_ab_synth.txt 2>&1; cat gpurun_out/r2o_ab_synth.txt
