cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cfg=${1:-trace4k}
python tools/one_frame.py $cfg 3 > gpurun_out/r3d_plain_$cfg.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r3d_prof_$cfg -f python tools/one_frame.py $cfg 3 > gpurun_out/r3d_ncu_$cfg.log 2>&1
cat gpurun_out/r3d_plain_$cfg.log
