"""
Problem definitions shared by the oracle tests (numpy) and the GPU parity tests.

Each returns plain numpy initial data in the reference's layout ``q[m,i(,j)]`` plus the
parameters the reference's test apps use:
  test/acoustics/1d/homogeneous/acoustics.py, test/acoustics/2d/homogeneous/acoustics.py,
  test/euler/2d/shockbubble.py, apps/shallow/2d/shallow2D.py (all under /root/reference).
"""
import numpy as np

GAMMA = 1.4
GAMMA1 = GAMMA - 1.


def centers(lower, upper, n):
    d = (upper - lower) / float(n)
    return np.array([lower + (i + 0.5) * d for i in range(n)]), d


def acoustics1d(mx=100):
    x, dx = centers(0.0, 1.0, mx)
    rho, bulk = 1.0, 1.0
    zz, cc = np.sqrt(rho * bulk), np.sqrt(rho / bulk)
    q = np.zeros((2, mx), order="F")
    beta, gamma, x0 = 100, 0, 0.75
    q[0, :] = np.exp(-beta * (x - x0) ** 2) * np.cos(gamma * (x - x0))
    return dict(q=q, d=[dx], params=[rho, bulk, cc, zz], dt_initial=dx / cc * 0.1,
                tfinal=1.0, nout=5)


def acoustics2d(mx=100, my=100, width=0.2):
    x, dx = centers(-1.0, 1.0, mx)
    y, dy = centers(-1.0, 1.0, my)
    Y, X = np.meshgrid(y, x)
    r = np.sqrt(X ** 2 + Y ** 2)
    q = np.zeros((3, mx, my), order="F")
    q[0] = (np.abs(r - 0.5) <= width) * (1. + np.cos(np.pi * (r - 0.5) / width))
    rho, bulk = 1.0, 4.0
    cc = np.sqrt(bulk / rho)
    zz = rho * cc
    return dict(q=q, d=[dx, dy], params=[rho, bulk, cc, zz],
                dt_initial=min(dx, dy) / cc * 0.45, tfinal=0.12, nout=10)


def shock_state(pinf=5.):
    rinf = (GAMMA1 + pinf * (GAMMA + 1.)) / ((GAMMA + 1.) + GAMMA1 * pinf)
    vinf = 1. / np.sqrt(GAMMA) * (pinf - 1.) / np.sqrt(0.5 * ((GAMMA + 1.) / GAMMA) * pinf + 0.5 * GAMMA1 / GAMMA)
    einf = 0.5 * rinf * vinf ** 2 + pinf / GAMMA1
    return rinf, vinf, einf


def shockbubble(mx=160, my=40, xupper=2.0, yupper=0.5, x0=0.5, y0=0., r0=0.2, rhoin=0.1):
    x, dx = centers(0.0, xupper, mx)
    y, dy = centers(0.0, yupper, my)
    Y, X = np.meshgrid(y, x)
    r = np.sqrt((X - x0) ** 2 + (Y - y0) ** 2)
    q = np.zeros((5, mx, my), order="F")
    q[0] = rhoin * (r <= r0) + 1. * (r > r0)
    q[3] = (1. * (r <= r0) + 1. * (r > r0)) / GAMMA1
    q[4] = 1. * (r <= r0)
    aux = np.zeros((1, mx, my), order="F")
    for j, yc in enumerate(y):
        aux[0, :, j] = yc
    return dict(q=q, aux=aux, d=[dx, dy], params=[GAMMA, GAMMA1], dt_initial=0.005,
                tfinal=0.2, nout=1, limiters=[4, 4, 4, 4, 2])


def shockbc_numpy(idim, t, qbc, mbc):
    """test/euler/2d/shockbubble.py:41-57 on a numpy qbc."""
    rinf, vinf, einf = shock_state()
    for i in range(mbc):
        qbc[0, i, ...] = rinf
        qbc[1, i, ...] = rinf * vinf
        qbc[2, i, ...] = 0.
        qbc[3, i, ...] = einf
        qbc[4, i, ...] = 0.


def _sdiv(xp, c, arr):
    """scalar / array with one IEEE division per element.  (torch evaluates
    ``python_float / tensor`` as ``tensor.reciprocal() * python_float``, which rounds twice.)"""
    if xp is np:
        return c / arr
    return xp.div(xp.full_like(arr, c), arr)


def euler_rad_src(xp, q, aux, dt):
    """test/euler/2d/shockbubble.py:59-94; ``xp`` is numpy or torch (same op order)."""
    dt2 = dt / 2.
    ndim = 2
    rad = aux[0]
    rho = q[0]
    u = q[1] / rho
    v = q[2] / rho
    press = GAMMA1 * (q[3] - 0.5 * rho * (u * u + v * v))
    qstar = xp.empty_like(q)
    c2 = _sdiv(xp, dt2 * (ndim - 1), rad)
    qstar[0] = q[0] - c2 * q[2]
    qstar[1] = q[1] - c2 * rho * u * v
    qstar[2] = q[2] - c2 * rho * v * v
    qstar[3] = q[3] - c2 * v * (q[3] + press)
    rho = qstar[0]
    u = qstar[1] / rho
    v = qstar[2] / rho
    press = GAMMA1 * (qstar[3] - 0.5 * rho * (u * u + v * v))
    c1 = _sdiv(xp, dt * (ndim - 1), rad)
    q[0] = q[0] - c1 * qstar[2]
    q[1] = q[1] - c1 * rho * u * v
    q[2] = q[2] - c1 * rho * v * v
    q[3] = q[3] - c1 * v * (qstar[3] + press)


def shallow2d(mx=150, my=150, rad=0.5, hl=2., hr=1.):
    x, dx = centers(-2.5, 2.5, mx)
    y, dy = centers(-2.5, 2.5, my)
    Y, X = np.meshgrid(y, x)
    r = np.sqrt(X ** 2 + Y ** 2)
    q = np.zeros((3, mx, my), order="F")
    q[0] = hl * (r <= rad) + hr * (r > rad)
    return dict(q=q, d=[dx, dy], params=[1.0], tfinal=2.5, nout=10)


def random_state(rp, shape, seed=0):
    """Seeded, physically admissible random data for kernel-level parity tests."""
    rng = np.random.RandomState(seed)
    if rp == "acoustics":
        meqn = 2 if len(shape) == 1 else 3
        q = rng.uniform(-1, 1, (meqn,) + tuple(shape))
    elif rp == "advection":
        q = rng.uniform(0, 1, (1,) + tuple(shape))
    elif rp == "euler":
        q = np.empty((5,) + tuple(shape))
        rho = rng.uniform(0.2, 2.0, shape)
        u = rng.uniform(-1.5, 1.5, shape)
        v = rng.uniform(-1.5, 1.5, shape)
        p = rng.uniform(0.2, 3.0, shape)
        q[0], q[1], q[2] = rho, rho * u, rho * v
        q[3] = p / GAMMA1 + 0.5 * rho * (u * u + v * v)
        q[4] = rng.uniform(0, 1, shape)
    elif rp == "shallow":
        q = np.empty((3,) + tuple(shape))
        h = rng.uniform(0.5, 2.0, shape)
        q[0], q[1], q[2] = h, h * rng.uniform(-2.0, 2.0, shape), h * rng.uniform(-2.0, 2.0, shape)
    else:
        raise ValueError(rp)
    return np.asfortranarray(q)


def smooth_state(rp, shape, seed=0):
    """Seeded smooth-plus-jump data (WENO reconstructions of white noise leave the
    admissible set for Euler); still exercises every entropy-fix branch."""
    rng = np.random.RandomState(seed)
    grids = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")

    def field(lo, hi):
        f = np.zeros(shape)
        for _ in range(3):
            ph = rng.uniform(0, 2 * np.pi, len(shape))
            k = rng.randint(1, 4, len(shape))
            term = np.ones(shape)
            for g, kk, p in zip(grids, k, ph):
                term = term * np.sin(2 * np.pi * kk * g + p)
            f += term / 3.0
        cut = rng.uniform(0.3, 0.7)
        f += 0.5 * (grids[0] > cut) - 0.25
        f += 0.01 * rng.uniform(-1, 1, shape)
        f = (f - f.min()) / (f.max() - f.min())
        return lo + (hi - lo) * f

    if rp == "acoustics":
        meqn = 2 if len(shape) == 1 else 3
        q = np.stack([field(-1, 1) for _ in range(meqn)])
    elif rp == "advection":
        q = np.stack([field(0, 1)])
    elif rp == "euler":
        rho, u, v, p = field(0.6, 1.6), field(-1.6, 1.6), field(-1.6, 1.6), field(0.6, 1.6)
        q = np.stack([rho, rho * u, rho * v, p / GAMMA1 + 0.5 * rho * (u * u + v * v), field(0, 1)])
    elif rp == "shallow":
        h = field(0.6, 1.6)
        q = np.stack([h, h * field(-1.6, 1.6), h * field(-1.6, 1.6)])
    else:
        raise ValueError(rp)
    return np.asfortranarray(q)


# ---------------------------------------------------------------------------
# shallow water on the sphere (test/shallow_sphere/shallow_4_Rossby_Haurwitz_wave.py)
# ---------------------------------------------------------------------------
SPHERE_G = 11489.57219


def sphere_problem(mx=40, my=20, mbc=2):
    """Initial data through the oracle's restatement of setaux.f / qinit.f (host, init time)."""
    from oracle import pyclaw_oracle as po
    xlower, xupper, ylower, yupper = -3.0, 1.0, -1.0, 1.0
    dx, dy = (xupper - xlower) / mx, (yupper - ylower) / my
    auxtmp = po.sphere_setaux(mbc, mx, my, xlower, ylower, dx, dy)
    qtmp = po.sphere_qinit(mbc, mx, my, xlower, ylower, dx, dy)
    return dict(q=np.asfortranarray(qtmp[:, mbc:-mbc, mbc:-mbc]), aux=np.asfortranarray(auxtmp[:, mbc:-mbc, mbc:-mbc]),
                auxbc_full=auxtmp, d=[dx, dy], lower=[xlower, ylower], params=[SPHERE_G], tfinal=10.0, nout=10)


def sphere_qbc_lower_y(idim, t, qbc, mbc):
    """shallow_4_Rossby_Haurwitz_wave.py:292-300 (idim is 1 here)"""
    for j in range(mbc):
        qbc1D = qbc[:, :, 2 * mbc - 1 - j].copy()
        qbc[:, :, j] = qbc1D[:, ::-1]


def sphere_qbc_upper_y(idim, t, qbc, mbc):
    my = qbc.shape[2] - 2 * mbc
    for j in range(mbc):
        qbc1D = qbc[:, :, my + mbc - 1 - j].copy()
        qbc[:, :, my + mbc + j] = qbc1D[:, ::-1]


# ---------------------------------------------------------------------------
# BASELINE-size inputs: separable (outer-product) fields, cheap to build at 8192^2
# ---------------------------------------------------------------------------
def big_state(rp, shape, seed=0):
    """Seeded smooth-plus-jump data built from 1-D profiles, f(i, j) = a(i) + b(j) + c(i) d(j):
    a 67-M-cell field costs a few outer products instead of a meshgrid of transcendentals.
    Velocities span both signs and exceed the sound speed in places, so every entropy-fix
    branch, both limiter branches (s > 0, s < 0) and zero / non-zero jumps are exercised."""
    rng = np.random.RandomState(seed)
    nx, ny = shape
    x = np.linspace(0.0, 1.0, nx)
    y = np.linspace(0.0, 1.0, ny)

    def prof(t):
        p = np.zeros_like(t)
        for _ in range(3):
            p += np.sin(2 * np.pi * rng.randint(1, 6) * t + rng.uniform(0, 2 * np.pi)) / 3.0
        p += 0.6 * (t > rng.uniform(0.2, 0.8)) - 0.3      # one jump
        k = rng.randint(0, len(t) - 8)
        p[k:k + 8] = p[k]                                 # a flat patch: zero jumps
        return p

    def field(lo, hi):
        f = np.add.outer(prof(x), prof(y)) + np.multiply.outer(prof(x), prof(y))
        fmin, fmax = f.min(), f.max()
        f -= fmin
        f *= (hi - lo) / (fmax - fmin)
        f += lo
        return f

    if rp == "acoustics":
        q = np.empty((3, nx, ny), order="F")
        for m in range(3):
            q[m] = field(-1, 1)
    elif rp == "euler":
        q = np.empty((5, nx, ny), order="F")
        rho, u, v, p = field(0.6, 1.6), field(-1.6, 1.6), field(-1.6, 1.6), field(0.6, 1.6)
        q[0] = rho
        q[1] = rho * u
        q[2] = rho * v
        q[3] = p / GAMMA1 + 0.5 * rho * (u * u + v * v)
        del u, v, p
        q[4] = field(0, 1)
    elif rp == "shallow":
        q = np.empty((3, nx, ny), order="F")
        h = field(0.6, 1.6)
        q[0] = h
        q[1] = h * field(-1.6, 1.6)
        q[2] = h * field(-1.6, 1.6)
    else:
        raise ValueError(rp)
    return q
