cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( for i in 1 2; do
  echo "RR_ROW_PROFILE=0"; RR_ROW_PROFILE=0 python tools/ab_kernel.py --cfg=synth4k,march4k,trace4k --reps=12 default 2>/dev/null | head -3
  echo "RR_ROW_PROFILE=1"; RR_ROW_PROFILE=1 python tools/ab_kernel.py --cfg=synth4k,march4k,trace4k --reps=12 default 2>/dev/null | head -3
done ) > gpurun_out/r3k_ab_row_profile.txt 2>&1
cat gpurun_out/r3k_ab_row_profile.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "bvh or placed or parity" > gpurun_out/r3k_pytest_subset.log 2>&1; tail -2 gpurun_out/r3k_pytest_subset.log
