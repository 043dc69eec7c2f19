# two GPUs: the in-library multi-GPU context (bit-identical to one GPU), then bench.py strong (default, with weak_value) under torchrun
cd $GRAFT_REPO_ROOT
nvidia-smi -L; free -g | head -2
timeout 600 python -m pytest tests/test_cluster_gpu.py -m gpu -x -q -k "multi_gpu or cluster_blocks" 2>&1 | tail -3
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 2 ) > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 600 gpurun_out/r2_bench_2gpu.err
python - <<PY
import json
l=json.loads([x for x in open("gpurun_out/r2_bench_2gpu.json").read().strip().splitlines() if x.startswith("{")][-1])
e=l["e2e"]
print("N=2 value %.3e (%.0f ms)" % (l["value"], l["ms_per_step"]), "weak %.3e" % l["config"]["weak_value"], "e2e %.3e (%.0f ms, blocks %d)" % (e["value"], e["ms_per_step"], e["row_blocks"]), l["config"]["phases_s"])
print(l["stats_last_step"])
PY
