# new pool-stage defaults on the full-size resident pass and on the chain-bound job (old defaults and a more aggressive setting beside them), then the full bench
cd $GRAFT_REPO_ROOT
timeout 400 python tools/knob_bench.py 2332800 10 "" "SPLASH_POOL_STAGE1=8,SPLASH_POOL_STAGE2=128,SPLASH_POOL_LANES=16,SPLASH_POOL_CTAS=48" "SPLASH_POOL_STAGE1=2,SPLASH_POOL_STAGE2=16,SPLASH_POOL_LANES=4,SPLASH_POOL_CTAS=192" "" 2>&1 | grep -v Warning | tee gpurun_out/r2_pool_full.log
timeout 400 python tools/knob_bench.py 583200 2 "" "SPLASH_POOL_STAGE1=2,SPLASH_POOL_STAGE2=16,SPLASH_POOL_LANES=4,SPLASH_POOL_CTAS=192" "SPLASH_POOL_STAGE1=2,SPLASH_POOL_STAGE2=16,SPLASH_POOL_LANES=8,SPLASH_POOL_CTAS=96" "SPLASH_POOL_STAGE1=4,SPLASH_POOL_STAGE2=32,SPLASH_POOL_LANES=4,SPLASH_POOL_CTAS=192" 2>&1 | grep -v Warning | tee gpurun_out/r2_pool_small.log
( time timeout 900 python bench.py --steps 3 --warmup 1 ) > gpurun_out/r2_bench_full4.json 2> gpurun_out/r2_bench_full4.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_bench_full4.json").read().strip().splitlines()[-1])
e=l["e2e"]
print("value %.3e (%.0f ms)" % (l["value"], l["ms_per_step"]), "bulk alone %.3e frac %.3f" % (l["roofline"]["cell_days_per_s"], l["roofline"]["frac"]), "e2e %.3e (%.0f ms, blocks %d)" % (e["value"], e["ms_per_step"], e["row_blocks"]), l["config"]["phases_s"])
print([ (round(b["total_ms"]), round(b["pool_wait_ms"])) for b in e["block_stats_last_step"]])
PY
