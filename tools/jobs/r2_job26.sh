cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_edge_gpu.py tests/test_special_cells_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/knob_bench.py 583200 2 "" "" 2>&1 | grep -v Warning | tee gpurun_out/r2_final_chain2.log
timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep -v Warning | tee -a gpurun_out/r2_final_chain2.log
