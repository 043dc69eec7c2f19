# last pool stage branch-light + stage 4 for the cells that keep declining it: chain-bound job, whole pass (timeline)
cd $GRAFT_REPO_ROOT
timeout 300 python tools/knob_bench.py 583200 2 "SPLASH_CHAIN_FAST_STAGES=4" "SPLASH_CHAIN_FAST_STAGES=0" 2>&1 | grep -v Warning | tee gpurun_out/r2_chain_fast3.log
SPLASH_TRACE=1 timeout 500 python tools/knob_bench.py 2332800 10 "SPLASH_CHAIN_FAST_STAGES=4" "SPLASH_CHAIN_FAST_STAGES=0" > gpurun_out/r2_trace_fast2.log 2>&1
grep -v "splash trace\|Warning" gpurun_out/r2_trace_fast2.log | tail -3
