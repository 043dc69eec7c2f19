# shipped configuration: tile priorities A/B, level table, full bench, ncu captures (bulk kernel + launch list), config-5 smoke
cd $GRAFT_REPO_ROOT
timeout 300 python tools/knob_bench.py 2332800 10 "" "SPLASH_TILE_PRIO=0" "" 2>&1 | grep -v Warning | tee gpurun_out/r2_prio.log
( time timeout 900 python bench.py --steps 3 --warmup 1 ) > gpurun_out/r2_bench_full3.json 2> gpurun_out/r2_bench_full3.err; echo "bench rc=$?"
timeout 900 python tools/level_table.py screen > gpurun_out/level_screen.log 2>&1; echo "screen rc=$?"
for v in level0 level1; do SPLASH_CUDA_LIB=$PWD/build/variants/libsplash_$v.so timeout 300 python tools/level_table.py run $v 2>&1 | tail -1; done
timeout 600 python tools/config5_smoke.py > gpurun_out/r2_config5.json 2> gpurun_out/r2_config5.err; echo "config5 rc=$?"; tail -c 1500 gpurun_out/r2_config5.json; tail -3 gpurun_out/r2_config5.err
# ncu: only after the plain command has exited 0
bash tools/gpu_job_profile.sh
python bench.py --cells 291600 --years 2 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_bench_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench_cmd.csv python bench.py --cells 291600 --years 2 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench_small.log 2>&1
echo "launch list rc=$?"; ls -la gpurun_out | tail -20; du -sh gpurun_out
