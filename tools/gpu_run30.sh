cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/ab_kernel.py --cfg=march4k --reps=12 ab/lib_rot.so ab/lib_rot0.so ab/lib_rot8.so ab/lib_rot30.so > gpurun_out/r3h_ab_march_rot2.txt 2>&1
cat gpurun_out/r3h_ab_march_rot2.txt
