# ncu capture of the last pool stage (k_pool_spin) with and without day_state_fast
cd $GRAFT_REPO_ROOT
timeout 200 python tools/chain_latency.py 2048 60 > gpurun_out/plain_chain.log 2>&1; echo "plain rc=$?"; grep pool_cells gpurun_out/plain_chain.log
for k in 4 0; do
SPLASH_CHAIN_FAST_STAGES=$k timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_pool_spin --launch-skip 2 -c 1 -o gpurun_out/prof_chain_f$k -f python tools/chain_latency.py 2048 60 > gpurun_out/ncu_chain_f$k.log 2>&1
echo "capture $k rc=$?"; tail -2 gpurun_out/ncu_chain_f$k.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
