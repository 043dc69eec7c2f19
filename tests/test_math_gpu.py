"""The device's own exp / log / acos / sin (csrc/splash_math.cuh) against 200-bit references.

They replace the libm calls of the reference's day step; the parity budget needs them at the few-ulp
level (glibc: < 1 ulp, CUDA libdevice: 1-2 ulp)."""
import mpmath as mp
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
mp.mp.prec = 200


def max_ulp(got, fn, x):
    worst = 0.0
    for g, xx in zip(got, x):
        r = fn(mp.mpf(float(xx)))
        rd = float(r)
        u = np.spacing(abs(rd)) if rd != 0 else 5e-324
        worst = max(worst, float(abs(mp.mpf(float(g)) - r) / mp.mpf(float(u))))
    return worst


@pytest.mark.parametrize("op,fn,lo,hi", [
    ("exp", mp.exp, -60.0, 60.0), ("exp", mp.exp, -1.0, 1.0), ("exp", mp.exp, -690.0, 690.0),
    ("log", mp.log, 1e-300, 1e300), ("log", mp.log, 0.5, 2.0), ("log", mp.log, 1e-6, 1.0),
    ("acos", mp.acos, -1.0, 1.0), ("acos", mp.acos, -0.5, 0.5),
    ("sin", mp.sin, 0.0, np.pi), ("sin", mp.sin, 0.0, 1e-3),
])
def test_accuracy(ctx, op, fn, lo, hi):
    rng = np.random.default_rng(hash((op, lo, hi)) % 2**32)
    if op == "log" and hi / max(lo, 1e-300) > 1e6:
        x = np.exp(rng.uniform(np.log(lo), np.log(hi), 4000))
    else:
        x = rng.uniform(lo, hi, 4000)
    got = ctx.debug_math(op, x)
    assert max_ulp(got, fn, x) <= 2.0


def test_special_values_follow_libdevice(ctx):
    assert np.array_equal(ctx.debug_math("exp", [np.inf, -np.inf, 800.0, -800.0, 0.0]), [np.inf, 0.0, np.inf, 0.0, 1.0])
    assert np.isnan(ctx.debug_math("exp", [np.nan])[0])
    r = ctx.debug_math("log", [0.0, -1.0, np.inf, 1.0, np.nan, 5e-324])
    assert r[0] == -np.inf and np.isnan(r[1]) and r[2] == np.inf and r[3] == 0.0 and np.isnan(r[4])
    assert abs(r[5] - np.log(5e-324)) < 1e-12
    r = ctx.debug_math("acos", [1.0, -1.0, 1.5, np.nan, 0.0])
    assert r[0] == 0.0 and r[1] == np.pi and np.isnan(r[2]) and np.isnan(r[3]) and r[4] == np.pi / 2
    r = ctx.debug_math("sin", [0.0, np.pi, np.nan, -0.3, 7.0])
    assert r[0] == 0.0 and abs(r[1] - 1.2246467991473532e-16) < 1e-30 and np.isnan(r[2])
    assert abs(r[3] - np.sin(-0.3)) < 1e-15 and abs(r[4] - np.sin(7.0)) < 1e-15


def test_branch_free_bodies_equal_what_they_replace(ctx):
    """day_state_fast (csrc/splash_model.cuh) calls the guard-free bodies: inside their guards they must give the bits
    of the guarded functions, and sqrt / division the bits of the compiler's IEEE expansions."""
    rng = np.random.default_rng(5)
    n = 1 << 21
    same = lambda a, b: np.array_equal(a, b, equal_nan=True)
    x = np.concatenate([rng.uniform(-1.0, 1.0, n), np.nextafter(1.0, 0.0) * np.sign(rng.uniform(-1, 1, 64)), rng.uniform(-0.5, 0.5, n // 4) ** 3])
    assert same(ctx.debug_math("acos_body", x), ctx.debug_math("acos", x))
    x = np.concatenate([np.exp(rng.uniform(-600.0, 600.0, n)), rng.uniform(0.0, 4.0, n), 1.0 - rng.uniform(0.0, 1.0, n // 4) ** 8])
    a, b = ctx.debug_math("sqrt_body", x), ctx.debug_math("sqrt", x)
    assert same(a, b) and np.array_equal(b, np.sqrt(x))
    x = np.concatenate([np.exp(rng.uniform(-650.0, 650.0, n)) * np.sign(rng.uniform(-1, 1, n)), rng.uniform(0.0, 1.0, n),
                        rng.integers(1, 1 << 20, n).astype(np.float64)])
    a, b = ctx.debug_math("div_body", x), ctx.debug_math("div", x)
    assert same(a, b) and not np.isnan(a).all()
    w = x[(np.arange(x.size, dtype=np.int64) * 7919 + 13) % x.size]
    ok = ~np.isnan(b)
    assert np.array_equal(b[ok], (x / w)[ok])
    x = rng.uniform(-699.0, 699.0, n)
    assert same(ctx.debug_math("exp_body", x), ctx.debug_math("exp", x))
    x = np.exp(rng.uniform(-700.0, 700.0, n))
    assert same(ctx.debug_math("log_body", x), ctx.debug_math("log", x))
