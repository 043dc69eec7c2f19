"""cells outside the parity gates on the 12 000-cell synthetic draw of the GPU parity test (seed 11, one
year) for the literal-order build (SPLASH_LEVEL=0), the shipped level-1 build, and the reference's own -mfma build.
usage (GPU box):  python tools/level_table.py screen        # CPU: reference run + conditioning screen, cached in gpurun_out/
                  SPLASH_CUDA_LIB=<variant.so> python tools/level_table.py run <name>   # one GPU build; appends to the table
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import _abi, api
from tests import conditioning, parity
from tests import oracle_lib as ol
from tests.synthetic import make_problem

CACHE = "/tmp/level_table_screen.npz"   # (hundreds of MB: stays on the box)
OUT = "gpurun_out/r02_level_table.json"
prob, dates = make_problem(n_cells=12000, n_years=1, seed=11)


def gates_off(got, ref):
    dev = conditioning.cell_deviation(got, ref)
    off = np.zeros(prob.n_cells, dtype=bool)
    for k, d in dev.items():
        off |= ~(d <= (1e-9 if k in parity.FLUX else (1e-8 if k == "sm_lim" else 1e-6)))
    return off


if sys.argv[1] == "screen":
    ref = ol.run_checked(prob, monthly=False)
    sparse, knocked = conditioning.stable_cells(prob, ref, conditioning.SPARSE)
    dense, knocked_d = conditioning.stable_cells(prob, ref, conditioning.DENSE)
    fma_off = gates_off(conditioning.run_variant("FMA", prob), ref)       # the reference built with -mfma -ffp-contract=fast, full gates
    np.savez_compressed(CACHE, sparse=sparse, dense=dense & sparse, fma_off=fma_off, **{"ref_" + k: ref[k] for k in _abi.OUTPUT_NAMES},
                        ref_diag=ref["cell_diag"])
    table = {"problem": "tests/synthetic.make_problem(12000 cells, 1 year, seed 11): the draw of test_live_oracle_agrees_at_scale",
             "checked_against_compiled_reference": bool(ref["checked_against_ref"]),
             "reference_unstable_sparse": int((~sparse).sum()), "densely_stable": int((dense & sparse).sum()), "knocked_out": {**knocked, **knocked_d},
             "reference_built_with_mfma": {"outside_gates": int(fma_off.sum()), "of_which_sparsely_stable": int((fma_off & sparse).sum()),
                                           "of_which_densely_stable": int((fma_off & dense & sparse).sum())}, "gpu": {}}
    json.dump(table, open(OUT, "w"), indent=1)
    print(json.dumps(table))
else:
    name = sys.argv[2]
    z = np.load(CACHE)
    ref = {k: z["ref_" + k] for k in _abi.OUTPUT_NAMES}
    ref["cell_diag"] = z["ref_diag"]
    got = api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au, prob.resolution, dates,
                          monthly_out=False, return_diag=True, tile_cells=4096)
    off = gates_off(got, ref)
    ip = _abi.DIAG_NAMES.index("spin_passes")
    row = {"outside_gates": int(off.sum()), "of_which_sparsely_stable": int((off & z["sparse"]).sum()),
           "of_which_densely_stable": int((off & z["dense"]).sum()), "also_outside_in_mfma_reference": int((off & z["fma_off"]).sum()),
           "nan_masks_equal": bool(all(np.array_equal(np.isnan(got[k]), np.isnan(ref[k])) for k in _abi.OUTPUT_NAMES)),
           "spin_passes_differ": int((got["cell_diag"][ip] != ref["cell_diag"][ip]).sum()), "lib": os.environ.get("SPLASH_CUDA_LIB", "default")}
    table = json.load(open(OUT))
    table["gpu"][name] = row
    json.dump(table, open(OUT, "w"), indent=1)
    print(name, json.dumps(row))
