CMD="python bench.py --cells 291600 --years 2 --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "rc=$?"; cut -c1-300 gpurun_out/bench_small.json; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
