cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/ab_kernel.py --cfg=synth4k --reps=12 ab/lib_cur.so ab/lib_st10.so ab/lib_st12.so ab/lib_st16.so > gpurun_out/r3p_ab_bvh_stack.txt 2>&1
cat gpurun_out/r3p_ab_bvh_stack.txt
