cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r3j_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3j_pytest_gpu.log
tail -3 gpurun_out/r3j_pytest_gpu.log
( for i in 1 2; do
  echo "RR_MARCH_PROFILE=0"; RR_MARCH_PROFILE=0 python tools/ab_kernel.py --cfg=march4k --reps=12 default | head -1
  echo "RR_MARCH_PROFILE=1"; RR_MARCH_PROFILE=1 python tools/ab_kernel.py --cfg=march4k --reps=12 default | head -1
done ) > gpurun_out/r3j_ab_march_profile.txt 2>&1
cat gpurun_out/r3j_ab_march_profile.txt
