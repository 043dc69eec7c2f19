"""How sensitive is the reference algorithm itself to the last bit of libm?  (evidence tool, CPU only)

Runs the C restatement (bit-identical to the compiled reference, tests/test_oracle_cpu.py) on seeded
synthetic cells three more times with a deliberately perturbed math library:
  POW     1 pow() call in 8 returns the neighbouring double (+-1 ulp)
  EXPLOG  1 exp()/log() call in 8 returns the neighbouring double
  FMA     same source built with -mfma -ffp-contract=fast (what -march=native would give R users)
and reports, per variant, how many cells leave the parity gates (|d wn| > 1e-6 mm anywhere, different
spin-up pass count).  These cells are ill-conditioned in the reference: no implementation on another
libm (libdevice included) can match them to the gates.  Usage: libm_sensitivity.py [cells] [years] [seed]
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rsplash_b200 import _abi  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

OUT = os.path.join(ROOT, "build", "sens")
PERTURB_H = """#include <math.h>
double pp_pow(double, double); double pp_exp(double); double pp_log(double);
#ifdef PERTURB_POW
#define pow pp_pow
#endif
#ifdef PERTURB_EXPLOG
#define exp pp_exp
#define log pp_log
#endif
"""
PERTURB_C = """#include <math.h>
#include <stdint.h>
#include <string.h>
static inline uint64_t mix(uint64_t x){x^=x>>33;x*=0xff51afd7ed558ccdULL;x^=x>>33;x*=0xc4ceb9fe1a85ec53ULL;x^=x>>33;return x;}
static inline double bump(double r, uint64_t h){
    if (!isfinite(r) || r==0.0 || (h & 7) != 0) return r;   /* 1 call in 8 */
    return nextafter(r, (h & 8) ? INFINITY : -INFINITY);
}
/* constant exponents GCC folds into multiplies/sqrt in the reference build are left alone */
double pp_pow(double x, double y){ uint64_t a,b; memcpy(&a,&x,8); memcpy(&b,&y,8);
    if (y==2.0||y==-1.0||y==0.5||y==3.0) return pow(x,y); return bump(pow(x,y), mix(a^mix(b))); }
double pp_exp(double x){ uint64_t a; memcpy(&a,&x,8); return bump(exp(x), mix(a+1)); }
double pp_log(double x){ uint64_t a; memcpy(&a,&x,8); return bump(log(x), mix(a+2)); }
"""


def build_variants():
    os.makedirs(OUT, exist_ok=True)
    open(os.path.join(OUT, "perturb.h"), "w").write(PERTURB_H)
    open(os.path.join(OUT, "perturb.c"), "w").write(PERTURB_C)
    src = os.path.join(ROOT, "oracle", "splash_oracle.c")
    base = ["gcc", "-std=gnu11", "-O2", "-fPIC", "-pthread", "-shared", "-I" + os.path.join(ROOT, "oracle")]
    libs = {}
    # perturb.c is compiled on its own: it must see the real libm names, not the macros
    pobj = os.path.join(OUT, "perturb.o")
    subprocess.run(["gcc", "-std=gnu11", "-O2", "-fPIC", "-fno-builtin", "-c", "-o", pobj, os.path.join(OUT, "perturb.c")], check=True)
    for v in ("POW", "EXPLOG"):
        so = os.path.join(OUT, f"libsplash_oracle_{v}.so")
        subprocess.run(base + ["-ffp-contract=off", f"-DPERTURB_{v}", "-include", os.path.join(OUT, "perturb.h"), "-o", so,
                               src, pobj, "-lm"], check=True)
        libs[v] = so
    so = os.path.join(OUT, "libsplash_oracle_FMA.so")
    subprocess.run(base + ["-mfma", "-ffp-contract=fast", "-o", so, src, "-lm"], check=True)
    libs["FMA"] = so
    return libs


def run_with(so, prob):
    lib = C.CDLL(so)
    lib.splash_oracle_grid_run.argtypes = [C.POINTER(_abi.SplashGridIn), C.POINTER(_abi.SplashOpts),
                                           C.POINTER(_abi.SplashGridOut), C.c_int]
    cout, arrays = ol.alloc_out(prob.n_days, prob.n_cells)
    opts = _abi.SplashOpts()
    cin = prob.c_in()
    assert lib.splash_oracle_grid_run(C.byref(cin), C.byref(opts), C.byref(cout), 0) == 0
    return arrays


def main():
    n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
    n_years = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 11
    prob, _ = make_problem(n_cells, n_years, seed=seed)
    base = ol.run_cpu(prob, monthly=False, core="oracle")
    ip = _abi.DIAG_NAMES.index("spin_passes")
    res = {"n_cells": n_cells, "n_days": prob.n_days, "seed": seed, "variants": {}}
    for name, so in build_variants().items():
        got = run_with(so, prob)
        r = {"passes_equal": int((got["cell_diag"][ip] == base["cell_diag"][ip]).sum())}
        for k in ("wn", "ro", "snow", "pet", "aet", "netr"):
            ok = np.isfinite(got[k]) & np.isfinite(base[k])
            dd = np.where(ok, np.abs(got[k] - base[k]), 0.0).max(0)
            r[k] = {"nan_mask_equal": bool(np.array_equal(np.isnan(got[k]), np.isnan(base[k]))),
                    "cells_abs_gt_1e-6": int((dd > 1e-6).sum()), "max_abs": float(dd.max())}
        res["variants"][name] = r
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
