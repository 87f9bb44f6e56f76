cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3q_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3q_pytest_gpu.log
tail -3 gpurun_out/r3q_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
