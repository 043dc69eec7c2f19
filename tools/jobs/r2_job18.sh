# (a) chain with the dry-regime closed forms; (b) uniform kernels on the branch-light state half (build variants)
cd $GRAFT_REPO_ROOT
timeout 300 python tools/knob_bench.py 583200 2 "SPLASH_CHAIN_FAST=1" "SPLASH_CHAIN_FAST=1,SPLASH_POOL_LANES=16" "SPLASH_CHAIN_FAST=1,SPLASH_POOL_LANES=32,SPLASH_POOL_CTAS=48" 2>&1 | grep -v Warning | tee gpurun_out/r2_chain_fast5.log
for v in uf uf128; do
  echo "== $v"
  SPLASH_CUDA_LIB=$PWD/build/variants/libsplash_$v.so timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep -v Warning | tail -1
done 2>&1 | tee gpurun_out/r2_uf.log
echo "== default"; timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep -v Warning | tail -1 | tee -a gpurun_out/r2_uf.log
