cd $GRAFT_REPO_ROOT
timeout 300 python tools/knob_bench.py 291600 10 "" "SPLASH_ROUNDS_RT=12,SPLASH_POOL_CAP=40000" "SPLASH_ROUNDS_RT=8,SPLASH_POOL_CAP=40000" "SPLASH_ROUNDS_RT=5,SPLASH_POOL_CAP=60000" 2>&1 | grep -v Warning | tee gpurun_out/r2_rounds_small.log
