# raw next-day forcing prefetch: whole pass + bulk alone; straggler-chain knobs on a chain-bound job; GPU tests; config-5 with 4 lanes
cd $GRAFT_REPO_ROOT
timeout 300 python tools/knob_bench.py 2332800 10 "" "" 2>&1 | grep -v Warning | tee gpurun_out/r2_prefetch.log
timeout 600 python tools/knob_bench.py 583200 2 "" "SPLASH_POOL_LANES=8,SPLASH_POOL_CTAS=96" "SPLASH_POOL_LANES=4,SPLASH_POOL_CTAS=192" "SPLASH_POOL_LANES=2,SPLASH_POOL_CTAS=384" "SPLASH_POOL_STAGE2=48" "SPLASH_POOL_STAGE2=48,SPLASH_POOL_LANES=4,SPLASH_POOL_CTAS=192" "SPLASH_POOL_STAGE1=4,SPLASH_POOL_STAGE2=32,SPLASH_POOL_LANES=8,SPLASH_POOL_CTAS=96" 2>&1 | grep -v Warning | tee gpurun_out/r2_chain_knobs.log
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests3.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_tests3.log
timeout 600 python tools/config5_smoke.py 4000 4000 2 32 4 > gpurun_out/r2_config5_l4.json 2> gpurun_out/r2_config5_l4.err; echo "config5 rc=$?"; tail -c 900 gpurun_out/r2_config5_l4.json
