# final validation of the round: GPU tests, bench.py with the driver's arguments, reference arm, launch list of a bench command
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests6.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_tests6.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r02c_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r02c_bench_1gpu.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02c_bench_reference_arm.json 2> gpurun_out/r02c_bench_reference_arm.err; echo "ref rc=$?"
python bench.py --cells 291600 --years 2 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_bench_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02c_launches_bench_cmd.csv python bench.py --cells 291600 --years 2 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench_small.log 2>&1
echo "launch list rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/r02c_bench_1gpu.json").read().strip().splitlines()[-1])
e=l["e2e"]
print("value %.3e (%.0f ms)" % (l["value"], l["ms_per_step"]), "bulk alone %.3e frac %.3f" % (l["roofline"]["cell_days_per_s"], l["roofline"]["frac"]), "e2e %.3e (%.0f ms, blocks %d)" % (e["value"], e["ms_per_step"], e["row_blocks"]), l["config"]["phases_s"], l["clocks"])
print({k: l["parity"][k] for k in ("cells_compared","nan_masks_equal","cells_outside_gates","outside_of_which_reference_stable_sparse")})
print([(round(b["total_ms"]),round(b["h2d_ms"]),round(b["pool_wait_ms"])) for b in e["block_stats_last_step"]])
r=json.loads(open("gpurun_out/r02c_bench_reference_arm.json").read().strip().splitlines()[-1]); print("reference arm %.3e" % r["value"], r["ms_per_step"])
PY
