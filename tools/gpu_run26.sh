cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/ab_kernel.py --cfg=trace4k,trace8k --reps=25 "$@" > gpurun_out/r3c_ab.txt 2>&1
cat gpurun_out/r3c_ab.txt
