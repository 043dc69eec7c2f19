# chain kernel with day_state_fast per pool stage: chain-bound job, chain latency, whole resident pass
cd $GRAFT_REPO_ROOT
for k in 0 4 6 7; do
  echo "== SPLASH_CHAIN_FAST_STAGES=$k"
  SPLASH_CHAIN_FAST_STAGES=$k timeout 200 python tools/chain_latency.py 4096 300 2>&1 | grep "pool_cells" | tail -1
done 2>&1 | tee gpurun_out/r2_chain_fast.log
timeout 300 python tools/knob_bench.py 583200 2 "SPLASH_CHAIN_FAST_STAGES=0" "SPLASH_CHAIN_FAST_STAGES=4" "SPLASH_CHAIN_FAST_STAGES=6" "SPLASH_CHAIN_FAST_STAGES=7" 2>&1 | grep -v Warning | tee -a gpurun_out/r2_chain_fast.log
timeout 400 python tools/knob_bench.py 2332800 10 "SPLASH_CHAIN_FAST_STAGES=0" "SPLASH_CHAIN_FAST_STAGES=4" "SPLASH_CHAIN_FAST_STAGES=6" 2>&1 | grep -v Warning | tee -a gpurun_out/r2_chain_fast.log
