for V in default e2 e4 t384 t256; do
  if [ $V = default ]; then unset SPLASH_CUDA_LIB; else export SPLASH_CUDA_LIB=$PWD/build/libsplash_$V.so; fi
  for i in 1 2; do python tools/profile_case.py bulk 151552 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V', 'bulk_ms %.2f' % d['bulk_ms'])"; done
done
unset SPLASH_CUDA_LIB
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/profile_case.py full 6144 1 > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
