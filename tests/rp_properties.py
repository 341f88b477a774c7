"""
Oracle-independent properties of the Riemann solvers.

Most solvers on BASELINE's configurations live in the external clawpack/riemann repository
(SURVEY.md 8(c), appendix B): `rpt2_acoustics`, `rp*_advection`, `rpn2/rpt2_shallow_roe_with_efix`
have no golden file in the reference, and `rpt2_euler_5wave` is in-tree but its results are
pinned by none.  The CUDA kernels agree with the oracle bit for bit, but the oracle restates the
same recollection -- so these checks tie BOTH to the mathematics the solvers are defined by
(LeVeque 2002, ch. 15, 21):

  W  sum_w wave_w          = q_r - q_l                      (the waves decompose the jump)
  C  amdq + apdq           = f(q_r) - f(q_l)                (Roe solvers are conservative; holds
                                                            with the entropy fix, which only moves
                                                            part of a wave's flux between A- and A+)
  S  amdq + apdq           = sum_w s_w wave_w               (fluctuations are built from the waves)
  E  A(q^) wave_w          = s_w wave_w                     (each wave is an eigenvector of the Roe
                                                            matrix with eigenvalue s_w)
  T  bmasdq + bpasdq       = B(q^) asdq                     (the transverse solver splits asdq with the
                                                            Jacobian of the OTHER direction at the
                                                            same Roe state)
  U  bmasdq carries only negative speeds, bpasdq only positive ones
                                                            (checked through  B^- = (B - |B|)/2 :
                                                            bmasdq = B^-(q^) asdq, bpasdq = B^+(q^) asdq)

`solve(ixy, ql, qr)` and `transverse(ixy, ql, qr, imp, asdq)` are supplied by the caller: the
oracle on the CPU (tests/test_rp_properties.py) or the CUDA kernels through the C ABI
(tests/test_gpu_rp_properties.py).
"""
import numpy as np

GAMMA, GAMMA1 = 1.4, 0.4
GRAV = 1.3


# ---------------------------------------------------------------------------------------------
# seeded states, including transonic rarefactions (entropy fix active), supersonic flow in both
# directions, strong jumps, identical states and states at rest
# ---------------------------------------------------------------------------------------------
def euler_states(n, seed):
    rng = np.random.RandomState(seed)

    def one():
        rho = rng.uniform(0.1, 3.0, n)
        u = rng.uniform(-3.0, 3.0, n)
        v = rng.uniform(-3.0, 3.0, n)
        p = rng.uniform(0.1, 4.0, n)
        psi = rng.uniform(0.0, 1.0, n)
        return np.stack([rho, rho * u, rho * v, p / GAMMA1 + 0.5 * rho * (u * u + v * v), psi])
    ql, qr = one(), one()
    k = n // 8
    qr[:, :k] = ql[:, :k]                       # zero jump
    ql[1:3, k:2 * k] = 0.0                      # gas at rest on the left
    qr[1:3, k:2 * k] = 0.0
    # transonic rarefaction in the normal direction: u_l - c_l < 0 < u_r - c_r
    sl = slice(2 * k, 3 * k)
    rho, p = 1.0, 1.0
    c = np.sqrt(GAMMA * p / rho)
    for q, un in ((ql, 0.5 * c), (qr, 1.8 * c)):
        q[0, sl] = rho
        q[3, sl] = p / GAMMA1 + 0.5 * rho * (un * un)
        q[4, sl] = 0.5
    return ql, qr, (sl, 0.5 * c, 1.8 * c)


def shallow_states(n, seed):
    rng = np.random.RandomState(seed)

    def one():
        h = rng.uniform(0.1, 3.0, n)
        return np.stack([h, h * rng.uniform(-3.0, 3.0, n), h * rng.uniform(-3.0, 3.0, n)])
    ql, qr = one(), one()
    k = n // 8
    qr[:, :k] = ql[:, :k]
    ql[1:, k:2 * k] = 0.0
    qr[1:, k:2 * k] = 0.0
    return ql, qr


def _set_normal(q, ixy, sl, un, rho):
    mu = 1 if ixy == 1 else 2
    mv = 2 if ixy == 1 else 1
    q[mu, sl] = rho * un
    q[mv, sl] = 0.0


# ---------------------------------------------------------------------------------------------
# fluxes and Jacobians
# ---------------------------------------------------------------------------------------------
def euler_flux(q, ixy):
    rho, mx_, my_, e, psi = q
    u, v = mx_ / rho, my_ / rho
    p = GAMMA1 * (e - 0.5 * rho * (u * u + v * v))
    un = u if ixy == 1 else v
    f = np.stack([rho * un, mx_ * un, my_ * un, (e + p) * un, psi * un])
    f[1 if ixy == 1 else 2] += p
    return f


def euler_roe(ql, qr):
    """Roe averages as the solver defines them (rpn2_euler_5wave.f:87-104)."""
    sl, sr = np.sqrt(ql[0]), np.sqrt(qr[0])
    pl = GAMMA1 * (ql[3] - 0.5 * (ql[1] ** 2 + ql[2] ** 2) / ql[0])
    pr = GAMMA1 * (qr[3] - 0.5 * (qr[1] ** 2 + qr[2] ** 2) / qr[0])
    u = (ql[1] / sl + qr[1] / sr) / (sl + sr)
    v = (ql[2] / sl + qr[2] / sr) / (sl + sr)
    H = ((ql[3] + pl) / sl + (qr[3] + pr) / sr) / (sl + sr)
    return u, v, H


def euler_jacobian(u, v, H, ixy):
    """d f / d q (ixy = 1) or d g / d q (ixy = 2) of the 4 Euler equations plus the tracer that
    is advected with the normal velocity (the 5-wave solver's 5th equation in its
    quasi-linear form: the 5th wave carries the jump in q5 with speed u).  Shape [n, 5, 5]."""
    n = len(u)
    g1 = GAMMA1
    q2 = u * u + v * v
    A = np.zeros((n, 5, 5))
    if ixy == 1:
        un, a, b = u, 1, 2
        ua, ub = u, v
    else:
        un, a, b = v, 2, 1
        ua, ub = v, u
    A[:, 0, a] = 1.0
    A[:, a, 0] = 0.5 * g1 * q2 - ua * ua
    A[:, a, a] = (3.0 - GAMMA) * ua
    A[:, a, b] = -g1 * ub
    A[:, a, 3] = g1
    A[:, b, 0] = -ua * ub
    A[:, b, a] = ub
    A[:, b, b] = ua
    A[:, 3, 0] = ua * (0.5 * g1 * q2 - H)
    A[:, 3, a] = H - g1 * ua * ua
    A[:, 3, b] = -g1 * ua * ub
    A[:, 3, 3] = GAMMA * ua
    A[:, 4, 4] = un
    return A


def shallow_flux(q, ixy):
    h, hu, hv = q
    u, v = hu / h, hv / h
    un = u if ixy == 1 else v
    f = np.stack([h * un, hu * un, hv * un])
    f[1 if ixy == 1 else 2] += 0.5 * GRAV * h * h
    return f


def shallow_roe(ql, qr):
    sl, sr = np.sqrt(ql[0]), np.sqrt(qr[0])
    u = (ql[1] / sl + qr[1] / sr) / (sl + sr)
    v = (ql[2] / sl + qr[2] / sr) / (sl + sr)
    c2 = GRAV * 0.5 * (ql[0] + qr[0])
    return u, v, c2


def shallow_jacobian(u, v, c2, ixy):
    n = len(u)
    A = np.zeros((n, 3, 3))
    if ixy == 1:
        a, b, ua, ub = 1, 2, u, v
    else:
        a, b, ua, ub = 2, 1, v, u
    A[:, 0, a] = 1.0
    A[:, a, 0] = c2 - ua * ua
    A[:, a, a] = 2.0 * ua
    A[:, b, 0] = -ua * ub
    A[:, b, a] = ub
    A[:, b, b] = ua
    return A


def acoustics_jacobian(n, K, rho, ixy):
    A = np.zeros((n, 3, 3))
    a = 1 if ixy == 1 else 2
    A[:, 0, a] = K
    A[:, a, 0] = 1.0 / rho
    return A


def _split(B, lams):
    """B^-, B^+ of a batch of diagonalisable matrices whose DISTINCT eigenvalues are lams[k][n]:
    the spectral projectors are Lagrange polynomials in B, P_k = prod_{j != k} (B - l_j)/(l_k - l_j),
    so nothing but the Jacobian itself and its eigenvalues enters (no eigenvector formulas)."""
    n, m, _ = B.shape
    I = np.eye(m)[None]
    Bm, Bp = np.zeros_like(B), np.zeros_like(B)
    for k, lk in enumerate(lams):
        Pk = np.broadcast_to(I, B.shape).copy()
        for j, lj in enumerate(lams):
            if j != k:
                Pk = np.einsum('nij,njk->nik', Pk, (B - lj[:, None, None] * I) / (lk - lj)[:, None, None])
        Bm += np.minimum(lk, 0.0)[:, None, None] * Pk
        Bp += np.maximum(lk, 0.0)[:, None, None] * Pk
    return Bm, Bp


def _scale(*arrs):
    return max(1.0, max(float(np.abs(a).max()) for a in arrs))


# ---------------------------------------------------------------------------------------------
# the checks
# ---------------------------------------------------------------------------------------------
def check_normal(name, solve, ixy, ql, qr, flux, jac, tol=2e-12, ncons=None, wcomps=None):
    wave, s, amdq, apdq = solve(ixy, ql, qr)
    meqn, mw, n = wave.shape
    dq = qr - ql
    sc = _scale(ql, qr)
    wc = slice(None) if wcomps is None else wcomps
    errW = np.abs(wave.sum(axis=1) - dq)[wc].max() / sc
    assert errW < tol, "%s ixy=%d: sum of waves != jump (%g)" % (name, ixy, errW)
    sw = np.einsum('wn,mwn->mn', s, wave)
    errS = np.abs(amdq + apdq - sw).max() / _scale(sw)
    assert errS < tol, "%s ixy=%d: amdq + apdq != sum s*wave (%g)" % (name, ixy, errS)
    if flux is not None:
        df = (flux(qr, ixy) - flux(ql, ixy))[:ncons]
        errC = np.abs((amdq + apdq)[:ncons] - df).max() / _scale(df, flux(qr, ixy))
        assert errC < 50 * tol, "%s ixy=%d: not conservative (%g)" % (name, ixy, errC)
    if jac is not None:
        A = jac(ixy)
        Aw = np.einsum('nij,jwn->iwn', A, wave)
        errE = np.abs(Aw - s[None] * wave).max() / _scale(Aw)
        assert errE < 50 * tol, "%s ixy=%d: waves are not eigenvectors of the Roe matrix (%g)" % (name, ixy, errE)
    return wave, s, amdq, apdq


def check_transverse(name, transverse, ixy, ql, qr, asdq, jac_other, eig_other, tol=1e-10):
    B = jac_other(3 - ixy)
    Bm, Bp = _split(B, eig_other(3 - ixy))
    for imp in (1, 2):
        bm, bp = transverse(ixy, ql, qr, imp, asdq)
        want = np.einsum('nij,jn->in', B, asdq)
        sc = _scale(want)
        errT = np.abs(bm + bp - want).max() / sc
        assert errT < tol, "%s ixy=%d imp=%d: bm + bp != B(q^) asdq (%g)" % (name, ixy, imp, errT)
        wm = np.einsum('nij,jn->in', Bm, asdq)
        wp = np.einsum('nij,jn->in', Bp, asdq)
        errU = max(np.abs(bm - wm).max(), np.abs(bp - wp).max()) / sc
        assert errU < 100 * tol, "%s ixy=%d imp=%d: wrong up/down-wind split (%g)" % (name, ixy, imp, errU)


def run_all(solve_for, transverse_for, n=4096):
    """solve_for(name) / transverse_for(name) return the callables for one solver family."""
    # ---- Euler 5-wave Roe (rpn2 in tree and pinned by sb_density; rpt2 results unpinned) ----
    ql, qr, (sl, unl, unr) = euler_states(n, 1)
    for ixy in (1, 2):
        l, r = ql.copy(), qr.copy()
        _set_normal(l, ixy, sl, unl, 1.0)
        _set_normal(r, ixy, sl, unr, 1.0)
        u, v, H = euler_roe(l, r)
        jac = lambda d: euler_jacobian(u, v, H, d)
        a = np.sqrt(GAMMA1 * (H - 0.5 * (u * u + v * v)))
        eig = lambda d: [(u if d == 1 else v) - a, (u if d == 1 else v), (u if d == 1 else v) + a]
        # the 5th equation is the colour equation psi_t + u psi_x = 0 (rpn2_euler_5wave.f:159-163:
        # wave 5 carries the jump in q5 with speed u), not a conservation law: C holds for 1..4
        wave, s, amdq, apdq = check_normal("euler5", solve_for("euler"), ixy, l, r, euler_flux, jac, ncons=4)
        # the entropy fix must have acted on the transonic block: both fluctuations non-zero
        assert np.abs(amdq[0, sl]).min() > 0 and np.abs(apdq[0, sl]).min() > 0
        rng = np.random.RandomState(7 + ixy)
        asdq = rng.uniform(-1, 1, l.shape)
        check_transverse("euler5", transverse_for("euler"), ixy, l, r, asdq, jac, eig)
        check_transverse("euler5", transverse_for("euler"), ixy, l, r, amdq, jac, eig)
    # ---- shallow water Roe + entropy fix (RECALLED) ----
    ql, qr = shallow_states(n, 2)
    for ixy in (1, 2):
        u, v, c2 = shallow_roe(ql, qr)
        jac = lambda d: shallow_jacobian(u, v, c2, d)
        c = np.sqrt(c2)
        eig = lambda d: [(u if d == 1 else v) - c, (u if d == 1 else v), (u if d == 1 else v) + c]
        wave, s, amdq, apdq = check_normal("shallow", solve_for("shallow"), ixy, ql, qr, shallow_flux, jac)
        rng = np.random.RandomState(17 + ixy)
        check_transverse("shallow", transverse_for("shallow"), ixy, ql, qr, rng.uniform(-1, 1, ql.shape), jac, eig)
        check_transverse("shallow", transverse_for("shallow"), ixy, ql, qr, apdq, jac, eig)
    # ---- acoustics (rpn2 pinned by acoustics2D_solution; rpt2 RECALLED) ----
    rng = np.random.RandomState(3)
    ql, qr = rng.uniform(-1, 1, (3, n)), rng.uniform(-1, 1, (3, n))
    K, rho = 4.0, 1.0
    for ixy in (1, 2):
        jac = lambda d: acoustics_jacobian(n, K, rho, d)
        flux = lambda q, d: np.einsum('nij,jn->in', acoustics_jacobian(n, K, rho, d), q)
        # two waves for three equations: the jump in the transverse velocity is a stationary
        # wave (s = 0) that the solver does not store, so W holds for (p, normal velocity)
        wave, s, amdq, apdq = check_normal("acoustics", solve_for("acoustics"), ixy, ql, qr, flux, jac,
                                           wcomps=[0, 1 if ixy == 1 else 2])
        cc = np.full(n, np.sqrt(K / rho))
        eig = lambda d: [-cc, 0.0 * cc, cc]
        check_transverse("acoustics", transverse_for("acoustics"), ixy, ql, qr, rng.uniform(-1, 1, ql.shape), jac, eig)
    # ---- advection (RECALLED) ----
    ql, qr = rng.uniform(0, 1, (1, n)), rng.uniform(0, 1, (1, n))
    for uv in ((0.7, -0.4), (-0.3, 0.9)):
        for ixy in (1, 2):
            vel = lambda d: uv[d - 1]
            jac = lambda d: np.full((n, 1, 1), vel(d))
            flux = lambda q, d: vel(d) * q
            check_normal("advection", solve_for(("advection", uv)), ixy, ql, qr, flux, jac)
            eig = lambda d: [np.full(n, vel(d))]
            check_transverse("advection", transverse_for(("advection", uv)), ixy, ql, qr,
                             rng.uniform(-1, 1, ql.shape), jac, eig)
