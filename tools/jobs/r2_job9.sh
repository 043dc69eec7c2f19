# validation of the cp.async build: GPU tests, A/B of the bulk kernel, then bench.py with the driver's own arguments (time budget) and the reference arm
cd $GRAFT_REPO_ROOT
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests4.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_tests4.log
timeout 300 python tools/knob_bench.py 2332800 10 "" "" 2>&1 | grep -v Warning | tee gpurun_out/r2_cpasync.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_1gpu.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc=$?"; tail -3 gpurun_out/r02_bench_reference_arm.err
python - <<PY
import json
l=json.loads(open("gpurun_out/r02_bench_1gpu.json").read().strip().splitlines()[-1])
e=l["e2e"]
print("value %.3e (%.0f ms)" % (l["value"], l["ms_per_step"]), "bulk alone %.3e frac %.3f" % (l["roofline"]["cell_days_per_s"], l["roofline"]["frac"]), "e2e %.3e (%.0f ms, blocks %d)" % (e["value"], e["ms_per_step"], e["row_blocks"]), l["config"]["phases_s"], l["clocks"])
print(l["parity"])
r=json.loads(open("gpurun_out/r02_bench_reference_arm.json").read().strip().splitlines()[-1]); print("reference arm %.3e" % r["value"], r["ms_per_step"])
PY
