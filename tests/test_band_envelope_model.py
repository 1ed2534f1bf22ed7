"""Exactness of the band-restricted lower-envelope scan used by the ESDF far-field kernel (csrc/esdf.cu, K2e).

A pure-Python model of the kernel's stack rules against brute force, on random columns with ties, gaps and slopes:
  * Felzenszwalb's domination test by integer cross-multiplication (sdf_map.cpp:682-715 builds the same envelope with
    real-valued intersections),
  * a site is stored only if it beats the current top at X = hi,
  * a stack of ONE site is replaced by a site that is strictly better at X = lo,
  * the band's rows are answered by a pointer walk.
The model is the specification the CUDA code follows statement by statement; the GPU tests check the kernel itself
against the oracle."""
import random

INF = 10 ** 12


def brute(g, lo, hi):
    return [min([(X - x) ** 2 + g[x] ** 2 for x in range(len(g)) if g[x] is not None] or [INF]) for X in range(lo, hi + 1)]


def band_scan(g, lo, hi):
    sv, sf = [], []
    longest = 0
    for x, gx in enumerate(g):
        if gx is None:
            continue
        f = gx * gx + x * x
        while len(sv) >= 2 and (sf[-1] - sf[-2]) * (x - sv[-1]) >= (f - sf[-1]) * (sv[-1] - sv[-2]):
            sv.pop(); sf.pop()
        if sv and f - 2 * hi * x >= sf[-1] - 2 * hi * sv[-1]:
            continue                                    # loses to the top at X = hi
        if len(sv) == 1 and f - 2 * lo * x < sf[0] - 2 * lo * sv[0]:
            sv[0], sf[0] = x, f                         # the only site loses the whole band
            continue
        sv.append(x); sf.append(f)
        longest = max(longest, len(sv))
    out, k = [], 0
    for X in range(lo, hi + 1):
        if not sv:
            out.append(INF)
            continue
        ck = sf[k] - 2 * X * sv[k]
        while k < len(sv) - 1:
            cn = sf[k + 1] - 2 * X * sv[k + 1]
            if cn > ck:
                break
            ck, k = cn, k + 1
        out.append(ck + X * X)
    return out, longest


def random_column(rng):
    n = rng.randint(1, 48)
    mode = rng.random()
    if mode < 0.3:
        return [rng.choice([None, 0, 1, 2, 3]) for _ in range(n)]
    if mode < 0.6:
        return [rng.choice([None] + list(range(12))) for _ in range(n)]
    if mode < 0.8:
        a, b = rng.randint(-3, 3), rng.randint(0, 30)
        return [abs(a * x + b) + rng.choice([0, 0, 0, 1]) for x in range(n)]
    return [rng.choice([None, None, None, rng.randint(0, 40)]) for _ in range(n)]


def test_band_restricted_stack_is_exact_and_short():
    rng = random.Random(20261017)
    worst = 0
    for _ in range(30000):
        g = random_column(rng)
        lo = rng.randint(0, len(g) - 1)
        hi = rng.randint(lo, len(g) - 1)
        got, longest = band_scan(g, lo, hi)
        assert got == brute(g, lo, hi), (g, lo, hi)
        worst = max(worst, longest - (hi - lo + 1))
    assert worst <= 2        # the stack never holds more than (rows of the band) + 2 sites
