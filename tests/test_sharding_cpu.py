"""CPU suite: row sharding of the global grid and the multi-rank bookkeeping (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np

from rsplash_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_land_cells_sum_and_bounds():
    rows = synthetic.land_cells_per_row()
    assert rows.sum() == synthetic.N_CELLS_5ARCMIN and rows.max() <= synthetic.GRID_COLS and rows.min() >= 0
    lat = synthetic.row_latitudes()
    assert rows[lat < -56].sum() == 0


def test_shards_are_contiguous_balanced_and_cover_the_grid():
    rows = synthetic.land_cells_per_row()
    for world in (1, 2, 4, 8):
        sh = synthetic.shard_rows(rows, world)
        assert sh[0][0] == 0 and sh[-1][1] == rows.sum()
        assert all(sh[i][1] == sh[i + 1][0] for i in range(world - 1))
        sizes = np.array([b - a for a, b in sh])
        assert sizes.max() - sizes.min() <= 2 * synthetic.GRID_COLS  # balanced to within two rows
        cum = set(np.concatenate([[0], np.cumsum(rows)]).tolist())
        assert all(a in cum for a, _ in sh)  # shard boundaries fall on row boundaries


WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from rsplash_b200 import synthetic
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# bench.py's strong-scaling deal: blocks of grid rows round-robin to the ranks (the reference's block scheduler)
grid = synthetic.Grid(100000, bench.GRID_SEED)
idx = bench.seg_index(bench.rank_cells(grid, world, rank, True))
# every rank holds its own cells of the ONE grid: gather the index sets and check that they partition it
parts = [None] * world
dist.all_gather_object(parts, idx.tolist())
allidx = np.sort(np.concatenate([np.asarray(p, dtype=np.int64) for p in parts]))
assert np.array_equal(allidx, np.arange(100000)), "the ranks' cells do not partition the grid"
c0, c1 = 0, len(idx)
# the only cross-rank traffic of the path: job size (sum) and step time (max)
job = torch.tensor([float((c1 - c0) * 3652)], dtype=torch.float64)
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(job, op=dist.ReduceOp.SUM)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert job.item() == 100000 * 3652, job
assert t.item() == float(world)
if rank == 0:
    print("OK", c0, c1)
dist.destroy_process_group()
"""


def test_two_rank_gloo_bookkeeping(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29613", str(script)], capture_output=True, text=True, env=env,
                       timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_reference_arm_prints_contract_line():
    """bench.py --impl reference: the CPU arm of the contract, tiny sample so it runs in seconds."""
    import json

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "32", "--years", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "cell-days/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
