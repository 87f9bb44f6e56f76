set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in trace4k synth4k; do
  k=trace_kernel
  python tools/one_frame.py $cfg 3 > gpurun_out/r2v_plain_$cfg.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2v_prof_$cfg -f python tools/one_frame.py $cfg 3 > gpurun_out/r2v_ncu_$cfg.log 2>&1
  cat gpurun_out/r2v_plain_$cfg.log
done
ls -la gpurun_out | tail
