# A/B of kernel build variants on the full benchmark grid (bulk kernel alone + one whole pass), then a full bench.py run
cd $GRAFT_REPO_ROOT
for v in base noprefetch estrin inline estrin_inline regs128 nosync; do
  SPLASH_CUDA_LIB=$PWD/build/variants/libsplash_$v.so timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep -v Warning | sed "s/^(default)/$v/" | tee -a gpurun_out/r2_variants.log
done
( time timeout 900 python bench.py --steps 3 --warmup 1 ) > gpurun_out/r2_bench_full1.json 2> gpurun_out/r2_bench_full1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_full1.err
