# ncu capture of the final pool stage (branch-light route)
cd $GRAFT_REPO_ROOT
timeout 200 python tools/chain_latency.py 2048 60 > gpurun_out/plain_chain.log 2>&1; echo "plain rc=$?"; grep pool_cells gpurun_out/plain_chain.log
timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_pool_spin --launch-skip 3 -c 1 -o gpurun_out/prof_chain_final -f python tools/chain_latency.py 2048 60 > gpurun_out/ncu_chain_final.log 2>&1
echo "capture rc=$?"; tail -2 gpurun_out/ncu_chain_final.log | cut -c1-200
