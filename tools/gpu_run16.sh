cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for nb in 16 4; do
python tools/one_shard.py trace8k $nb 3 > gpurun_out/r2r_plain_shard$nb.log 2>&1 && timeout 600 ncu --set full --clock-control none -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r2r_prof_shard$nb -f python tools/one_shard.py trace8k $nb 3 > gpurun_out/r2r_ncu_shard$nb.log 2>&1
cat gpurun_out/r2r_plain_shard$nb.log
done
