# eight GPUs of one box: bench.py as the scaling driver launches it (strong scaling by default, weak_value beside it)
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l; free -g | head -2; nproc
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 2 ) > gpurun_out/r02c_bench_8gpu.json 2> gpurun_out/r02c_bench_8gpu.err; echo "bench8 rc=$?"
tail -c 800 gpurun_out/r02c_bench_8gpu.err
python - <<PY
import json
l=json.loads([x for x in open("gpurun_out/r02c_bench_8gpu.json").read().strip().splitlines() if x.startswith("{")][-1])
e=l["e2e"]
print("N=8 value %.3e (%.0f ms)" % (l["value"], l["ms_per_step"]), "weak %.3e" % l["config"]["weak_value"], "e2e %.3e (%.0f ms, blocks %d)" % (e["value"], e["ms_per_step"], e["row_blocks"]), l["config"]["phases_s"])
print(l["stats_last_step"])
PY
