# timeline of the resident pass with the old / branch-light last pool stage
cd $GRAFT_REPO_ROOT
SPLASH_TRACE=1 timeout 500 python tools/knob_bench.py 2332800 10 "SPLASH_CHAIN_FAST_STAGES=0" "SPLASH_CHAIN_FAST_STAGES=4" "SPLASH_CHAIN_FAST_STAGES=4,SPLASH_POOL_LANES=8" > gpurun_out/r2_trace_fast.log 2>&1
grep -v "splash trace\|Warning" gpurun_out/r2_trace_fast.log | tail -5
