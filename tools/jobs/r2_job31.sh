cd $GRAFT_REPO_ROOT
echo "== default"; timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep -v Warning | tail -1 | tee gpurun_out/r2_bulk_shapes3.log
for v in b736r88 b704r88 b672r96; do
  echo "== $v"
  SPLASH_CUDA_LIB=$PWD/build/variants/libsplash_$v.so timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep -v Warning | tail -1
done 2>&1 | tee -a gpurun_out/r2_bulk_shapes3.log
