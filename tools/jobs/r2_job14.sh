# chain kernel with the select-form day_state_fast (own sqrt / division bodies): math identities, chain-bound job, whole pass
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_math_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/knob_bench.py 583200 2 "SPLASH_CHAIN_FAST_STAGES=0" "SPLASH_CHAIN_FAST_STAGES=4" "SPLASH_CHAIN_FAST_STAGES=6" 2>&1 | grep -v Warning | tee gpurun_out/r2_chain_fast2.log
timeout 400 python tools/knob_bench.py 2332800 10 "SPLASH_CHAIN_FAST_STAGES=0" "SPLASH_CHAIN_FAST_STAGES=4" 2>&1 | grep -v Warning | tee -a gpurun_out/r2_chain_fast2.log
