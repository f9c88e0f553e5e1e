#!/usr/bin/env python
"""Minimax coefficients of P in  asin(sqrt t)/sqrt t = 1 + t*P(t)  on [0, tmax], weighted for the RELATIVE error of
asin (i.e. of f = 1 + t P): Remez exchange in 60-digit arithmetic (mpmath), printed as C hex-float literals.

    python tools/remez_asin.py DEGREE TMAX        e.g.  9 0.25   (the objective kernels' series)
"""
import sys

import mpmath as mp

mp.mp.dps = 60


def f(t):
    if t == 0:
        return mp.mpf(1)
    s = mp.sqrt(t)
    return mp.asin(s) / s


def g(t):                      # the function P approximates: (f - 1) / t
    if t == 0:
        return mp.mpf(1) / 6
    return (f(t) - 1) / t


def remez(deg, tmax, iters=40):
    n = deg + 2
    # Chebyshev extrema as the first reference
    xs = [tmax * (1 - mp.cos(mp.pi * i / (n - 1))) / 2 for i in range(n)]
    coef = None
    for _ in range(iters):
        # solve  P(x_i) + (-1)^i E * w(x_i) = g(x_i),  error of f is t*(P - g), relative: t*(P-g)/f  -> weight f/t
        A = mp.matrix(n, n)
        b = mp.matrix(n, 1)
        for i, x in enumerate(xs):
            for k in range(deg + 1):
                A[i, k] = x ** k
            wgt = f(x) / x if x != 0 else mp.mpf(10) ** 30      # at t = 0 the relative error is 0 whatever P(0) is
            A[i, deg + 1] = (-1) ** i * wgt
            b[i] = g(x)
        sol = mp.lu_solve(A, b)
        coef = [sol[k] for k in range(deg + 1)]
        E = sol[deg + 1]

        def err(t):
            if t == 0:
                return mp.mpf(0)
            p = mp.polyval(coef[::-1], t)
            return t * (p - g(t)) / f(t)
        # new extrema: dense scan + local refinement
        N = 4000
        grid = [tmax * i / N for i in range(N + 1)]
        vals = [err(t) for t in grid]
        ext = []
        for i in range(1, N):
            if (vals[i] - vals[i - 1]) * (vals[i + 1] - vals[i]) <= 0:
                ext.append(grid[i])
        ext = [grid[0]] + ext + [grid[-1]]
        # keep n alternating extrema with the largest magnitudes
        cand = sorted(ext, key=lambda t: -abs(err(t)))[: n + 2]
        cand = sorted(set(cand))
        if len(cand) < n:
            break
        # pick n points with alternating signs
        best = [cand[0]]
        for t in cand[1:]:
            if mp.sign(err(t)) == mp.sign(err(best[-1])):
                if abs(err(t)) > abs(err(best[-1])):
                    best[-1] = t
            else:
                best.append(t)
        if len(best) < n:
            break
        while len(best) > n:
            if abs(err(best[0])) < abs(err(best[-1])):
                best.pop(0)
            else:
                best.pop()
        if max(abs(a - b_) for a, b_ in zip(best, xs)) < mp.mpf(10) ** -25:
            xs = best
            break
        xs = best
    N = 20000
    emax = max(abs(err(tmax * i / N)) for i in range(N + 1))
    return coef, emax


if __name__ == "__main__":
    deg, tmax = int(sys.argv[1]), mp.mpf(sys.argv[2])
    coef, emax = remez(deg, tmax)
    print("degree %d on [0, %s]: max relative error of asin %.3e" % (deg, sys.argv[2], float(emax)))
    print(", ".join(float(c).hex() for c in coef))
