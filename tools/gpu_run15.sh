cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 300 python tools/shard_time.py trace8k 8; RR_PREWARM=1 timeout 300 python tools/shard_time.py trace8k 8; timeout 300 python tools/shard_time.py trace8k 4; RR_PREWARM=1 timeout 300 python tools/shard_time.py trace8k 4; timeout 300 python tools/shard_time.py trace8k 16; timeout 300 python tools/shard_time.py trace8k 2; timeout 300 python tools/shard_time.py trace4k 1; RR_PREWARM=1 timeout 300 python tools/shard_time.py trace4k 1) 2>&1 | grep static16 > gpurun_out/r2q_prewarm.txt; cat gpurun_out/r2q_prewarm.txt
