# bulk launch shapes: 768 x 80 / 768 x 84 / 640 x 96 with the forcing half's constants read from global memory
cd $GRAFT_REPO_ROOT
for v in b768 b768r84 b640; do
  echo "== $v"
  SPLASH_CUDA_LIB=$PWD/build/variants/libsplash_$v.so timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep -v Warning | tail -1
done 2>&1 | tee gpurun_out/r2_bulk_shapes.log
