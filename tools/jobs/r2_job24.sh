cd $GRAFT_REPO_ROOT
SPLASH_TRACE=1 timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep "declining cell" | sort | uniq | tee gpurun_out/r2_decliners.log | head -40
