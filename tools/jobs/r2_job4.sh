# register-cap variants (whole pass + bulk alone), the GPU test suite, the level-0 / level-1 table, one full bench
cd $GRAFT_REPO_ROOT
for v in i96 i104 i112 i120 i128; do
  SPLASH_CUDA_LIB=$PWD/build/variants/libsplash_$v.so timeout 300 python tools/knob_bench.py 2332800 10 "" 2>&1 | grep -v Warning | sed "s/^(default)/$v/" | tee -a gpurun_out/r2_variants2.log
done
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests2.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests2.log
timeout 900 python tools/level_table.py screen > gpurun_out/level_screen.log 2>&1; echo "screen rc=$?"
for v in level0 level1; do SPLASH_CUDA_LIB=$PWD/build/variants/libsplash_$v.so timeout 300 python tools/level_table.py run $v 2>&1 | tail -1; done
( time timeout 900 python bench.py --steps 3 --warmup 1 ) > gpurun_out/r2_bench_full2.json 2> gpurun_out/r2_bench_full2.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_full2.err
