cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_math_gpu.py -x -q 2>&1 | tail -2
SPLASH_TRACE=1 timeout 300 python tools/knob_bench.py 583200 2 "" 2> gpurun_out/r2_trace_small.err | grep -v Warning | tee gpurun_out/r2_final_chain.log
grep "declining cell" gpurun_out/r2_trace_small.err | sort | uniq | head
SPLASH_TRACE=1 timeout 300 python tools/knob_bench.py 2332800 10 "" 2> gpurun_out/r2_trace_big.err | grep -v Warning | tee -a gpurun_out/r2_final_chain.log
grep "declining cell" gpurun_out/r2_trace_big.err | sort | uniq | head
