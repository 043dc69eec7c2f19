cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_chain_fast_gpu.py tests/test_special_cells_gpu.py tests/test_math_gpu.py tests/test_edge_gpu.py -x -q 2>&1 | grep -E "passed|failed|error" | tail -3
