# 768-thread hybrid-constant shape for all uniform kernels: GPU test suite, A/B against the 512-thread build
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests5.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_tests5.log
timeout 300 python tools/knob_bench.py 2332800 10 "" "" 2>&1 | grep -v Warning | tee gpurun_out/r2_u768.log
SPLASH_CUDA_LIB=$PWD/build/variants/libsplash_u512.so timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep -v Warning | tee -a gpurun_out/r2_u768.log
timeout 300 python tools/knob_bench.py 583200 2 "" 2>&1 | grep -v Warning | tee -a gpurun_out/r2_u768.log
