"""How sensitive is the reference algorithm itself to the last bit of libm?  (evidence tool, CPU only)

Re-runs the C restatement (bit-identical to the compiled reference, tests/test_oracle_cpu.py) on seeded synthetic
cells against the perturbed math libraries of oracle/perturb/ (see tests/conditioning.py) and reports, per variant,
how many cells leave the parity gates of the unperturbed run.  Those cells are ill-conditioned in the reference:
no implementation on another libm (libdevice included) can match them to the gates.
usage: libm_sensitivity.py [cells] [years] [seed]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rsplash_b200 import _abi  # noqa: E402
from tests import conditioning  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402


def main():
    n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
    n_years = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 11
    prob, _ = make_problem(n_cells, n_years, seed=seed)
    base = ol.run_cpu(prob, monthly=False, core="oracle")
    ip = _abi.DIAG_NAMES.index("spin_passes")
    res = {"n_cells": n_cells, "n_days": prob.n_days, "seed": seed, "variants": {}}
    for v in conditioning.VARIANTS:
        got = conditioning.run_variant(v, prob)
        dev = conditioning.cell_deviation(got, base)
        r = {"passes_equal": int((got["cell_diag"][ip] == base["cell_diag"][ip]).sum())}
        for k in ("wn", "ro", "snow", "bflow"):
            r[k + "_cells_abs_gt_1e-6"] = int((~(dev[k] <= 1e-6)).sum())
        for k in ("pet", "aet", "netr", "cond"):
            r[k + "_cells_rel_gt_1e-9"] = int((~(dev[k] <= 1e-9)).sum())
        res["variants"][v] = r
    stable, knocked = conditioning.stable_cells(prob, base)
    res["stable_cells"] = int(stable.sum())
    res["knocked_out_at_a_tenth_of_the_gates"] = knocked
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
