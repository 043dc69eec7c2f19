"""Development aid (CPU): outputs of the host build of the device day step (levels 1 and 0) on a fixed set of cells,
saved or compared bit for bit.  Used when the day step is refactored WITHOUT changing its arithmetic (e.g. moving
literals into constant memory): `save` before, `check` after.
usage: emul_snapshot.py save|check [file]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import _abi
from tests import fixtures as fx
from tests import host_emul_harness as he
from tests.synthetic import make_problem

mode = sys.argv[1]
path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/emul_snapshot.npz"
out = {}
probs = {"bourne": fx.load_problem("bourne")[0], "syn": make_problem(1500, 2, seed=77)[0],
         "polar": make_problem(600, 1, seed=78, lat_range=(66.0, 89.0))[0], "tropic": make_problem(600, 1, seed=79, lat_range=(-20.0, 20.0))[0]}
for name, prob in probs.items():
    for level in (1, 0):
        r = he.run(prob, level=level)
        for k in _abi.OUTPUT_NAMES + ("state_final", "cell_diag"):
            out[f"{name}_l{level}_{k}"] = r[k]
if mode == "save":
    np.savez_compressed(path, **out)
    print("saved", len(out), "arrays to", path)
else:
    ref = np.load(path)
    bad = [k for k in out if not np.array_equal(out[k], ref[k], equal_nan=True)]
    print("bit-identical" if not bad else f"DIFFERENT: {bad[:8]} ({len(bad)} arrays)")
    sys.exit(1 if bad else 0)
