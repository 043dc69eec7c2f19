"""Evidence tool: the at-scale parity check of tests/test_parity_gpu.py on other seeds / sizes.
usage: parity_seeds.py <cells> <years> <seed> [<seed> ...]"""
import sys

sys.path.insert(0, ".")
from rsplash_b200 import api  # noqa: E402
from tests.test_parity_gpu import _check_synthetic  # noqa: E402

n_cells, n_years = int(sys.argv[1]), int(sys.argv[2])
ctx = api.default_context()
for seed in map(int, sys.argv[3:]):
    try:
        _check_synthetic(ctx, n_cells=n_cells, n_years=n_years, seed=seed, max_unstable=0.10)
        print("seed", seed, "ok", flush=True)
    except AssertionError as e:
        print("seed", seed, "FAILED:", str(e)[:300], flush=True)
