"""Coefficient tables of the WENO reconstructions of order 2k-1 (k = 3..9), regenerated from
their definition in exact rational arithmetic.

The reference carries them as ~2400 lines of PyWENO-generated Fortran
(src/fortran/1d/sharpclaw/weno.f90:5-2425, one subroutine per order).  Every subroutine has
the same shape (weno.f90:104-250 for k = 4):

    sigma_r = sum_{a<=b} S[r][a,b] q(i-r+a) q(i-r+b)            smoothness of stencil r
    omega_r = w[r] / (sigma_r + 1e-36)**2, normalised           once with the left-edge ideal
                                                                 weights, once with the right-edge
    f_r     = sum_j C[r][j] q(i-r+j)                            edge value from stencil r
    ql(i)   = sum_r omega^L_r f^L_r ,  qr(i) = sum_r omega^R_r f^R_r

with stencil r = cells i-r .. i-r+k-1.  The numbers below are those formulas' exact values:
reconstruction coefficients from the interpolation of the primitive function, ideal weights
from matching the (2k-1)-point reconstruction, smoothness indicators
sigma_r = sum_{l=1}^{k-1} int_cell h^(2l-1) (d^l p_r / dx^l)^2 dx  (Jiang & Shu 1996).
"""
from fractions import Fraction
from functools import lru_cache

import numpy as np


def _poly_mul(a, b):
    out = [Fraction(0)] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] += x * y
    return out


def _poly_der(a):
    return [a[i] * i for i in range(1, len(a))] or [Fraction(0)]


def _poly_int(a, lo, hi):
    tot = Fraction(0)
    for i, c in enumerate(a):
        tot += c * (hi ** (i + 1) - lo ** (i + 1)) / (i + 1)
    return tot


def _poly_eval(a, x):
    tot = Fraction(0)
    for c in reversed(a):
        tot = tot * x + c
    return tot


def _stencil_polys(k, r):
    """For stencil cells i-r .. i-r+k-1 (h = 1, cell i = [-1/2, 1/2]): the k polynomials
    P_j(x) such that p(x) = sum_j qbar_j P_j(x) has the given cell averages."""
    # edges of the stencil
    edges = [Fraction(-1, 2) - r + m for m in range(k + 1)]
    polys = []
    for j in range(k):
        # primitive V(x_edge[m]) = sum_{l<m} qbar_l ; with qbar = e_j: V = 0 for m<=j, 1 for m>j
        vals = [Fraction(1 if m > j else 0) for m in range(k + 1)]
        # Lagrange interpolation of V through the k+1 edges, then differentiate
        V = [Fraction(0)]
        for m in range(k + 1):
            if vals[m] == 0:
                continue
            num = [Fraction(1)]
            den = Fraction(1)
            for n in range(k + 1):
                if n != m:
                    num = _poly_mul(num, [-edges[n], Fraction(1)])
                    den *= edges[m] - edges[n]
            V = [a + b for a, b in zip(V + [Fraction(0)] * (len(num) - len(V)), [c * vals[m] / den for c in num])]
        polys.append(_poly_der(V))
    return polys


@lru_cache(maxsize=None)
def exact_tables(k):
    """Rational tables: CL[r][j], CR[r][j] (left / right edge), WL[r], WR[r], S[r][(a,b)]."""
    half = Fraction(1, 2)
    CL, CR, S = [], [], []
    for r in range(k):
        P = _stencil_polys(k, r)
        CL.append([_poly_eval(p, -half) for p in P])
        CR.append([_poly_eval(p, half) for p in P])
        quad = {}
        derivs = [list(p) for p in P]
        for l in range(1, k):
            derivs = [_poly_der(d) for d in derivs]
            for a in range(k):
                for b in range(a, k):
                    v = _poly_int(_poly_mul(derivs[a], derivs[b]), -half, half)
                    quad[(a, b)] = quad.get((a, b), Fraction(0)) + (v if a == b else 2 * v)
        S.append(quad)
    # ideal weights: the (2k-1)-cell reconstruction is the weighted sum of the k-cell ones
    big = _stencil_polys(2 * k - 1, k - 1)          # cells i-k+1 .. i+k-1
    W = []
    for edge, C in ((-half, CL), (half, CR)):
        target = [_poly_eval(p, edge) for p in big]   # coefficient of cell i-k+1+n
        w = [None] * k
        # stencil r is the only one containing cell i-r+k-1 ... solve from the extremes inwards:
        # cell i+k-1 (n = 2k-2) appears in stencil r = 0 only, cell i+k-2 in r = 0, 1, ...
        for r in range(k):
            n = 2 * k - 2 - r                        # cell i + k-1-r ; stencil r holds it at j = k-1
            acc = target[n]
            for rr in range(r):
                j = (k - 1 - r) + rr                 # position of that cell in stencil rr
                acc -= w[rr] * C[rr][j]
            w[r] = acc / C[r][k - 1]
        W.append(w)
    return {'CL': CL, 'CR': CR, 'WL': W[0], 'WR': W[1], 'S': S}


def tables(k, literals='f32'):
    """Flat float64 arrays for the kernels / the oracle.  ``literals='f32'`` rounds every
    coefficient (and 1e-36) to single precision first: the generated Fortran writes them
    without a kind suffix, so gfortran reads them as REAL(4)."""
    if k < 3 or k > 9:
        raise ValueError("weno_order must be an odd number between 5 and 17 (inclusive)")
    T = exact_tables(k)
    rnd = (lambda x: float(np.float32(float(x)))) if literals == 'f32' else float
    npair = k * (k + 1) // 2
    S = np.zeros((k, npair))
    for r in range(k):
        n = 0
        for a in range(k):
            for b in range(a, k):
                S[r, n] = rnd(T['S'][r][(a, b)])
                n += 1
    return {
        'k': k,
        'S': np.ascontiguousarray(S),
        'CL': np.array([[rnd(v) for v in row] for row in T['CL']]),
        'CR': np.array([[rnd(v) for v in row] for row in T['CR']]),
        'WL': np.array([rnd(v) for v in T['WL']]),
        'WR': np.array([rnd(v) for v in T['WR']]),
        'eps': rnd(1.0e-36),
    }
