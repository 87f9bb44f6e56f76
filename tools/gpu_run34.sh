set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r3m_smi.txt
nproc >> gpurun_out/r3m_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3m_pytest_gpu.log
tail -5 gpurun_out/r3m_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r3m_bench_n1.json 2> gpurun_out/r3m_bench_n1.err; echo "bench rc=$?"; tail -5 gpurun_out/r3m_bench_n1.err; head -c 1500 gpurun_out/r3m_bench_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r3m_bench_ref_n1.json 2> gpurun_out/r3m_bench_ref_n1.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3m_launches_bench_default4k.csv python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r3m_ncu_launches.log 2>&1
for cfg in trace4k synth4k march4k; do
  k=trace_kernel; [ $cfg = march4k ] && k=march_kernel
  python tools/one_frame.py $cfg 3 > gpurun_out/r3m_plain_$cfg.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r3m_prof_$cfg -f python tools/one_frame.py $cfg 3 > gpurun_out/r3m_ncu_$cfg.log 2>&1
  cat gpurun_out/r3m_plain_$cfg.log
done
python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" > gpurun_out/r3m_smoke.log 2>&1; tail -2 gpurun_out/r3m_smoke.log
timeout 900 python tools/parity_report.py --out gpurun_out/r3m_parity_report.md > gpurun_out/r3m_parity.log 2>&1; tail -3 gpurun_out/r3m_parity.log
