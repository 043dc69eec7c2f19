# A/B of the branch-light state half: chain latency (k_pool_spin) and the resident pass / bulk kernel per build variant
cd $GRAFT_REPO_ROOT
for v in base cf uf ufnofb; do
  echo "== $v"
  SPLASH_CUDA_LIB=$PWD/build/variants/libsplash_$v.so timeout 200 python tools/chain_latency.py 4096 300 2>&1 | grep -v Warning | grep "pool_cells\|Error\|error" | tail -3
  SPLASH_CUDA_LIB=$PWD/build/variants/libsplash_$v.so timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep -v Warning | tail -2
done 2>&1 | tee gpurun_out/r2_fast_ab.log
