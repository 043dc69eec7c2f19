# probe + two-role final pool stage: chain-bound job, whole pass (timeline)
cd $GRAFT_REPO_ROOT
timeout 300 python tools/knob_bench.py 583200 2 "SPLASH_CHAIN_FAST=1" "SPLASH_CHAIN_FAST=0" "SPLASH_CHAIN_FAST=1,SPLASH_POOL_LANES=4" 2>&1 | grep -v Warning | tee gpurun_out/r2_chain_fast4.log
SPLASH_TRACE=1 timeout 500 python tools/knob_bench.py 2332800 10 "SPLASH_CHAIN_FAST=1" "SPLASH_CHAIN_FAST=0" > gpurun_out/r2_trace_fast3.log 2>&1
grep -v "splash trace\|Warning" gpurun_out/r2_trace_fast3.log | tail -3
