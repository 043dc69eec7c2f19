# ncu capture of the 768-thread bulk kernel (two full waves), then bench.py with the driver's arguments and the reference arm
cd $GRAFT_REPO_ROOT
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum
python tools/profile_case.py bulk 227328 1 > gpurun_out/plain_bulk.log 2>&1 && \
ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_run_bulk -c 1 -o gpurun_out/prof_bulk -f python tools/profile_case.py bulk 227328 1 > gpurun_out/ncu_bulk.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_bulk.log | cut -c1-200
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02b_bench_1gpu.json 2> gpurun_out/r02b_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r02b_bench_1gpu.err
python - <<PY
import json
l=json.loads(open("gpurun_out/r02b_bench_1gpu.json").read().strip().splitlines()[-1])
e=l["e2e"]
print("value %.3e (%.0f ms)" % (l["value"], l["ms_per_step"]), "bulk alone %.3e frac %.3f" % (l["roofline"]["cell_days_per_s"], l["roofline"]["frac"]), "e2e %.3e (%.0f ms, blocks %d)" % (e["value"], e["ms_per_step"], e["row_blocks"]), l["config"]["phases_s"], l["clocks"])
print(l["parity"])
print(e["block_stats_last_step"])
PY
