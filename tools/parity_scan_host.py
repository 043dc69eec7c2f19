"""CPU-only parity scan (development / round planning): the device day step compiled for the host
(tests/host_emul) against the C restatement on many synthetic draws; reports the cells that are well-conditioned in
the reference (stable under every perturbed-libm variant) and still leave the gates -- candidates for places where
the level-1 arithmetic has to reproduce a floating-point accident of the reference.
usage: parity_scan_host.py n_cells n_years seed [seed ...]     (LAT_RANGE="50,72" restricts the latitudes, AU_LAYERS=1: scalar Au)"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import _abi  # noqa: E402
from tests import conditioning, parity  # noqa: E402
from tests import host_emul_harness as he  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

n_cells, n_years = int(sys.argv[1]), int(sys.argv[2])
for seed in map(int, sys.argv[3:]):
    kw = {"lat_range": tuple(float(v) for v in os.environ["LAT_RANGE"].split(","))} if os.environ.get("LAT_RANGE") else {}
    if os.environ.get("AU_LAYERS"):
        kw["au_layers"] = int(os.environ["AU_LAYERS"])
    prob, dates = make_problem(n_cells, n_years, seed=seed, **kw)
    ref = ol.run_cpu(prob, monthly=False, core="oracle")
    got = he.run(prob)
    sparse, rep = conditioning.stable_cells(prob, ref, conditioning.SPARSE)
    dense, rep2 = conditioning.stable_cells(prob, ref, conditioning.DENSE)
    off = np.zeros(n_cells, bool)
    for k, d in conditioning.cell_deviation(got, ref).items():
        off |= ~(d <= (1e-9 if k in parity.FLUX else (1e-8 if k == "sm_lim" else 1e-6)))
    a, b = got["cell_diag"], ref["cell_diag"]
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    doff = (np.nan_to_num(rel, nan=0.0, posinf=0.0) > 1e-12).any(0)
    nan_ok = all(np.array_equal(np.isnan(got[k]), np.isnan(ref[k])) for k in _abi.OUTPUT_NAMES)
    bad_s, bad_d = np.flatnonzero((off | doff) & sparse), np.flatnonzero((off | doff) & sparse & dense)
    print(f"seed {seed}: {n_cells}x{prob.n_days}  reference-unstable {int((~sparse).sum())}, densely stable {int((sparse & dense).sum())}; "
          f"host build outside the gates {int(off.sum())} (+{int((doff & ~off).sum())} diag only), of which sparsely stable {bad_s.tolist()}, "
          f"densely stable {bad_d.tolist()}; NaN masks equal {nan_ok}", flush=True)
