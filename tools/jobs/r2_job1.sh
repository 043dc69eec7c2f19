set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3; nproc; free -g | head -2
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests1.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests1.log
tail -15 gpurun_out/r2_tests1.log
( time timeout 600 python bench.py --cells 583200 --years 2 --steps 2 --warmup 1 ) > gpurun_out/r2_bench_small.json 2> gpurun_out/r2_bench_small.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_small.err
( time timeout 600 python tools/knob_bench.py 2332800 10 "" "SPLASH_REGIME_SORT=0" ) > gpurun_out/r2_knob1.log 2>&1; echo "knob rc=$?"
cat gpurun_out/r2_knob1.log | tail -8
