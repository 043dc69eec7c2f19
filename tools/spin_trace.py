"""Development aid: replay one cell's spin-up pass by pass through the resume entry, on the GPU and on the C
restatement from the SAME start state each pass, and report the first day of each pass where they part ways.
Usage: spin_trace.py n_cells n_years seed cell [cell ...]   (one-year problems: the series is the spin-up year)
SPIN_TRACE_HOST=1 replays through the host build of the device day step (tests/host_emul) instead of the GPU."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import _abi, api  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

n_cells, n_years, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prob, dates = make_problem(n_cells, n_years, seed=seed)
assert prob.n_days == 365
KEYS = ("wn", "snow", "ro", "aet", "pet", "cond", "bflow", "netr", "sm_lim")


def gpu(p, st):
    if os.environ.get("SPIN_TRACE_HOST"):
        from tests import host_emul_harness as he

        return he.run(p, state_init=st)
    return api.splash_grid(p.sw_in, p.tc, p.pn, p.lat, p.elev, p.slop, p.asp, p.soil, p.au, p.resolution, dates,
                           monthly_out=False, return_diag=True, return_state=True, **({} if st is None else {"state_init": st}))


def suspects():
    """cells stable under every perturbed-libm variant whose GPU outputs or float diagnostics leave the gates"""
    from tests import conditioning, parity
    ref = ol.run_cpu(prob, monthly=False, core="oracle")
    stable, _ = conditioning.stable_cells(prob, ref, conditioning.VARIANTS)
    got = gpu(prob, None)
    off = np.zeros(n_cells, bool)
    for k, d in conditioning.cell_deviation(got, ref).items():
        off |= ~(d <= (1e-9 if k in parity.FLUX else (1e-8 if k == "sm_lim" else 1e-6)))
    a, b = got["cell_diag"], ref["cell_diag"]
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    doff = (np.nan_to_num(rel, nan=0.0, posinf=0.0) > 1e-12).any(0)
    print("stable cells off in outputs:", np.flatnonzero(off & stable), " in diagnostics:", np.flatnonzero(doff & stable))
    return np.flatnonzero((off | doff) & stable)


cells = suspects() if sys.argv[4] == "auto" else map(int, sys.argv[4:])
for c in cells:
    c = int(c)
    p = prob.subset([c])
    full = ol.run_cpu(p, monthly=False, core="oracle")
    res = full["cell_diag"][_abi.DIAG_NAMES.index("RES"), 0]
    print(f"cell {c}: RES {res!r} AI {full['cell_diag'][_abi.DIAG_NAMES.index('AI'), 0]!r}")
    lateral = float(p.au[2, 0])
    for phase in (1, 2):
        st = np.array([[res], [0.0], [0.0], [0.0], [0.0], [lateral]])
        first_wn0 = None
        for k in range(1000):
            r = ol.run_cpu(p, monthly=False, core="oracle", state_init=st)
            g = gpu(p, st)
            worst = {key: float(np.nanmax(np.abs(g[key][:, 0] - r[key][:, 0]))) for key in KEYS}
            d = np.abs(g["wn"][:, 0] - r["wn"][:, 0])
            j = int(np.argmax(d > 1e-10)) if (d > 1e-10).any() else -1
            print(f"  spin-up {phase} pass {k}: max |d| " + " ".join(f"{a}={b:.1e}" for a, b in worst.items()) + f"  first day |dwn|>1e-10: {j}")
            if j >= 0:
                for day in range(max(0, j - 1), min(365, j + 3)):
                    print(f"      day {day} tc {p.tc[day, 0]!r} sw {p.sw_in[day, 0]!r} pn {p.pn[day, 0]!r}")
                    for key in KEYS:
                        print(f"         {key:7s} gpu {g[key][day, 0]!r} ref {r[key][day, 0]!r}")
            if k == 0 and phase == 1:
                ai = np.nansum(r["pet"][:, 0]) / np.nansum(p.pn[:, 0])
                print(f"    AI from pass 0: {ai!r}  (gpu pass: {np.nansum(g['pet'][:, 0]) / np.nansum(p.pn[:, 0])!r})")
            wn0 = r["wn"][0, 0]
            st = r["state_final"].copy()
            if first_wn0 is not None and abs(wn0 - prev_wn0) <= 1.0 and k >= 1:
                pass
            # convergence as SPLASH::spin_up does it: day 1 recomputed from the pass's last state vs the pass's day 1
            nxt = ol.run_cpu(p, monthly=False, core="oracle", state_init=st)["wn"][0, 0]
            prev_wn0 = first_wn0 = wn0
            if abs(nxt - wn0) <= 1.0:
                print(f"    converged after pass {k}: |{nxt!r} - {wn0!r}| <= 1")
                break
        lateral = ai
    print("  final spin state (replayed)", st[:, 0].tolist())
    print("  main run day 0 (full oracle) wn", full["wn"][0, 0])
