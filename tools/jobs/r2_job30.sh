cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests/test_chain_fast_gpu.py -x -q ) 2>&1 | tail -8
