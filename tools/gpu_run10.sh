set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/ab_kernel.py --cfg=march4k --reps=7 ab/lib_head.so default > gpurun_out/r2j_ab_march.txt 2>&1; cat gpurun_out/r2j_ab_march.txt
(timeout 300 python tools/shard_time.py trace8k 8; RR_NOFLUSH=1 timeout 300 python tools/shard_time.py trace8k 8; timeout 300 python tools/shard_time.py trace4k 1; RR_NOFLUSH=1 timeout 300 python tools/shard_time.py trace4k 1) 2>&1 | grep static16 > gpurun_out/r2j_flush.txt; cat gpurun_out/r2j_flush.txt
timeout 900 python tools/random_sweep.py 40 160 > gpurun_out/r2j_sweep.txt 2>&1; tail -2 gpurun_out/r2j_sweep.txt
for cfg in trace4k synth4k march4k; do
  k=trace_kernel; [ $cfg = march4k ] && k=march_kernel
  python tools/one_frame.py $cfg 3 > gpurun_out/r2j_plain_$cfg.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2j_prof_$cfg -f python tools/one_frame.py $cfg 3 > gpurun_out/r2j_ncu_$cfg.log 2>&1
  cat gpurun_out/r2j_plain_$cfg.log
done
