cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests7.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error" gpurun_out/r2_tests7.log | tail -3
