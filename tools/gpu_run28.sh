cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/ab_kernel.py --cfg=synth4k --reps=15 ab/lib_cur.so ab/lib_leaf3.so ab/lib_leaf2.so ab/lib_leaf2s10.so ab/lib_leaf1.so > gpurun_out/r3e_ab_leaf.txt 2>&1
cat gpurun_out/r3e_ab_leaf.txt
python tools/ab_e2e_geom.py > gpurun_out/r3e_e2e_geom.txt 2>&1
cat gpurun_out/r3e_e2e_geom.txt
bash tools/gpu_run27.sh trace4k
