cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/exp_flush_gap.py trace4k > gpurun_out/r2z_flush_gap.txt 2>&1
python tools/exp_flush_gap.py trace8k >> gpurun_out/r2z_flush_gap.txt 2>&1
cat gpurun_out/r2z_flush_gap.txt
