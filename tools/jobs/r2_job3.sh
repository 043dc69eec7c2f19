# whole bench.py (value + roofline + host-fed e2e with the rolling-window scheduler) under kernel build variants
cd $GRAFT_REPO_ROOT
for v in base inline s32_inline r128_inline r128_inline_nopf; do
  ( SPLASH_CUDA_LIB=$PWD/build/variants/libsplash_$v.so timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_var_$v.json 2> gpurun_out/r2_var_$v.err; echo "$v rc=$?" )
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2_var_$v.json").read().strip().splitlines()[-1])
e=l["e2e"]
print("$v", "value %.3e (%.0f ms)" % (l["value"], l["ms_per_step"]), "bulk alone %.3e" % l["roofline"]["cell_days_per_s"], "e2e %.3e (%.0f ms, refill %.1f s, blocks %d)" % (e["value"], e["ms_per_step"], e["caller_refill_s_inside_timed_region"], e["row_blocks"]), l["config"]["phases_s"])
PY
done
