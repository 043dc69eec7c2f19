cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests/test_chain_fast_gpu.py tests/test_parity_gpu.py tests/test_edge_gpu.py tests/test_cluster_gpu.py -x -q ) 2>&1 | tail -5
timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep -v Warning | tail -1 | tee gpurun_out/r2_final_pass.log
