cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/march_strip_times.py > gpurun_out/r3i_march_strips.txt 2>&1
head -3 gpurun_out/r3i_march_strips.txt; sort -t: -k2 -n -r gpurun_out/r3i_march_strips.txt | head -5
