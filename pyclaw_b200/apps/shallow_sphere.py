"""Shallow water on the sphere, 4-Rossby-Haurwitz wave: the application helpers of
/root/reference/apps/shallow-sphere (mapc2p.f, setaux.f, qinit.f, src2.f and the boundary
conditions of shallow_4_Rossby_Haurwitz_wave.py), for device-resident state.

* ``mapc2p``, ``setaux``, ``qinit`` run once at set-up, vectorised on the host in numpy (the
  reference runs them once through the f2py ``problem`` module).
* ``src2`` (Coriolis term, 4-stage Runge-Kutta, tangent-plane projection; called twice per
  step with Strang splitting) and the pole-fold boundary conditions are part of the time
  step and act on the device tensors.
"""
import numpy as np
import torch

Rsphere = 1.0
G_SW = 11489.57219          # classic2.sw.g in the reference script


def mapc2p(x1, y1, R=Rsphere):
    """mapc2p.f:2-76, vectorised.  Returns xp, yp, zp."""
    xc = np.array(x1, dtype=np.float64, copy=True)
    yc = np.array(y1, dtype=np.float64, copy=True)
    xc, yc = np.broadcast_arrays(xc, yc)
    xc, yc = xc.copy(), yc.copy()
    xc = np.where(xc >= 1.0, xc - 4.0, xc)
    xc = np.where(xc < -3.0, xc + 4.0, xc)
    m = yc >= 1.0
    yc = np.where(m, 2.0 - yc, yc)
    xc = np.where(m, -2.0 - xc, xc)
    m = yc < -1.0
    yc = np.where(m, -2.0 - yc, yc)
    xc = np.where(m, -2.0 - xc, xc)
    m = xc < -1.0
    xc = np.where(m, -2.0 - xc, xc)
    sgnz = np.where(m, -1.0, 1.0)
    sgnxc = np.where(np.signbit(xc), -1.0, 1.0)
    sgnyc = np.where(np.signbit(yc), -1.0, 1.0)
    xc1, yc1 = np.abs(xc), np.abs(yc)
    d = np.maximum(np.maximum(xc1, yc1), 1.e-10)
    DD = R * d * (2.0 - d) / np.sqrt(2.0)
    center = DD - np.sqrt(np.maximum(R ** 2 - DD ** 2, 0.0))
    xp = DD / d * xc1
    yp = DD / d * yc1
    up = yc1 > xc1
    yp = np.where(up, center + np.sqrt(np.maximum(R ** 2 - xp ** 2, 0.0)), yp)
    xp = np.where(up, xp, center + np.sqrt(np.maximum(R ** 2 - yp ** 2, 0.0)))
    zp = np.sqrt(np.maximum(R ** 2 - (xp ** 2 + yp ** 2), 0.0))
    return xp * sgnxc, yp * sgnyc, zp * sgnz


def setaux(mbc, mx, my, xlower, ylower, dxc, dyc, R=Rsphere):
    """setaux.f:2-218, vectorised: aux(16, mx+2mbc, my+2mbc) including ghost cells."""
    i = np.arange(1 - mbc, mx + mbc + 2)
    j = np.arange(1 - mbc, my + mbc + 2)
    xc = xlower + (i - 1.0) * dxc
    yc = ylower + (j - 1.0) * dyc
    xp, yp, zp = mapc2p(xc[:, None], yc[None, :], R)
    r = np.sqrt(xp ** 2 + yp ** 2)
    with np.errstate(invalid='ignore', divide='ignore'):
        theta = np.where(r > 1.e-4, np.arccos(np.clip(xp / np.where(r > 0, r, 1.0), -1.0, 1.0)), 0.0)
    theta = np.where(yp < 0.0, -theta, theta)
    ac = np.arccos(np.clip(r / R, -1.0, 1.0))
    phi = np.where(zp > 0.0, np.pi / 2.0 - ac, np.pi / 2.0 + ac)
    aux = np.empty((16, mx + 2 * mbc, my + 2 * mbc), order='F')
    c = (slice(0, -1), slice(0, -1))          # corner (i, j)
    cu = (slice(0, -1), slice(1, None))       # (i, j+1)
    cr = (slice(1, None), slice(0, -1))       # (i+1, j)
    cur = (slice(1, None), slice(1, None))    # (i+1, j+1)
    # left edge
    etx, ety, etz = xp[cu] - xp[c], yp[cu] - yp[c], zp[cu] - zp[c]
    aux[4], aux[5], aux[6] = etx, ety, etz
    erx, ery, erz = 0.5 * (xp[c] + xp[cu]), 0.5 * (yp[c] + yp[cu]), 0.5 * (zp[c] + zp[cu])
    enx, eny, enz = ety * erz - etz * ery, etz * erx - etx * erz, etx * ery - ety * erx
    nn = np.sqrt(enx ** 2 + eny ** 2 + enz ** 2)
    aux[1], aux[2], aux[3] = enx / nn, eny / nn, enz / nn
    # bottom edge
    etx, ety, etz = xp[cr] - xp[c], yp[cr] - yp[c], zp[cr] - zp[c]
    aux[10], aux[11], aux[12] = etx, ety, etz
    erx, ery, erz = 0.5 * (xp[c] + xp[cr]), 0.5 * (yp[c] + yp[cr]), 0.5 * (zp[c] + zp[cr])
    enx, eny, enz = ery * etz - erz * ety, erz * etx - erx * etz, erx * ety - ery * etx
    nn = np.sqrt(enx ** 2 + eny ** 2 + enz ** 2)
    aux[7], aux[8], aux[9] = enx / nn, eny / nn, enz / nn
    # radial vector at the cell centre
    ic = np.arange(1 - mbc, mx + mbc + 1)
    jc = np.arange(1 - mbc, my + mbc + 1)
    xpm, ypm, zpm = mapc2p((xlower + (ic - 0.5) * dxc)[:, None], (ylower + (jc - 0.5) * dyc)[None, :], R)
    aux[13], aux[14], aux[15] = xpm, ypm, zpm

    def beta(a, b):
        return np.sin(phi[a]) * np.sin(phi[b]) * np.cos(theta[a] - theta[b]) + np.cos(phi[a]) * np.cos(phi[b])
    acos = lambda v: np.arccos(np.clip(v, -1.0, 1.0))
    d12 = R * acos(beta(c, cr))
    d23 = R * acos(beta(cu, cr))
    d13 = R * acos(beta(cu, c))
    d24 = R * acos(beta(cur, cr))
    d34 = R * acos(beta(cur, cu))
    s123 = 0.5 * (d12 + d23 + d13)
    s234 = 0.5 * (d23 + d34 + d24)
    t123 = np.tan(s123 / 2.0) * np.tan((s123 - d12) / 2.0) * np.tan((s123 - d23) / 2.0) * np.tan((s123 - d13) / 2.0)
    t234 = np.tan(s234 / 2.0) * np.tan((s234 - d23) / 2.0) * np.tan((s234 - d34) / 2.0) * np.tan((s234 - d24) / 2.0)
    E123 = 4.0 * np.arctan(np.sqrt(np.maximum(t123, 0.0)))
    E234 = 4.0 * np.arctan(np.sqrt(np.maximum(t234, 0.0)))
    aux[0] = (E123 + E234) / (dxc * dyc)
    return aux


def qinit(mx, my, xlower, ylower, dx, dy, R=Rsphere):
    """qinit.f:3-107 (4-Rossby-Haurwitz wave), vectorised: q(4, mx, my)."""
    a, K, Omega, G, t0, h0, Rw = 6.37122e6, 7.848e-6, 7.292e-5, 9.80616, 86400.0, 8.e3, 4.0
    xc = xlower + (np.arange(1, mx + 1) - 0.5) * dx
    yc = ylower + (np.arange(1, my + 1) - 0.5) * dy
    xp, yp, zp = mapc2p(xc[:, None], yc[None, :], R)
    rad = np.maximum(np.sqrt(xp ** 2 + yp ** 2), 1.e-6)
    asn = np.arcsin(np.clip(np.abs(yp) / rad, -1.0, 1.0))
    theta = np.zeros_like(xp)
    theta = np.where((xp > 0) & (yp > 0), asn, theta)
    theta = np.where((xp < 0) & (yp > 0), np.pi - asn, theta)
    theta = np.where((xp < 0) & (yp < 0), -np.pi + asn, theta)
    theta = np.where((xp > 0) & (yp < 0), -asn, theta)
    phi = np.where(zp > 0, np.arcsin(np.clip(zp / R, -1, 1)), -np.arcsin(np.clip(-zp / R, -1, 1)))
    lam, cy, sy = theta, np.cos(phi), np.sin(phi)
    bigA = 0.5 * K * (2.0 * Omega + K) * cy ** 2.0 + 0.25 * K * K * cy ** (2.0 * Rw) * (
        (Rw + 1.0) * cy ** 2.0 + (2.0 * Rw * Rw - Rw - 2.0) - 2.0 * Rw * Rw * cy ** (-2.0))
    bigB = (2.0 * (Omega + K) * K) / ((Rw + 1.0) * (Rw + 2.0)) * cy ** Rw * (
        (Rw * Rw + 2.0 * Rw + 2.0) - (Rw + 1.0) ** 2 * cy ** 2)
    bigC = 0.25 * K * K * cy ** (2 * Rw) * ((Rw + 1.0) * cy ** 2 - (Rw + 2.0))
    Uin1 = (K * cy + K * cy ** (Rw - 1.) * (Rw * sy ** 2. - cy ** 2.) * np.cos(Rw * lam)) * t0
    Uin2 = (-K * Rw * cy ** (Rw - 1.) * sy * np.sin(Rw * lam)) * t0
    U1 = -np.sin(lam) * Uin1 - sy * np.cos(lam) * Uin2
    U2 = np.cos(lam) * Uin1 - sy * np.sin(lam) * Uin2
    U3 = cy * Uin2
    q = np.empty((4, mx, my), order='F')
    q[0] = h0 / a + (a / G) * (bigA + bigB * np.cos(Rw * lam) + bigC * np.cos(2.0 * Rw * lam))
    q[1], q[2], q[3] = q[0] * U1, q[0] * U2, q[0] * U3
    return q


_DF = float(np.float32(12.600576))   # "df=12.600576e0": a REAL(4) literal in src2.f:38


def src2(solver, state, dt):
    """src2.f:2-147 as one fused kernel (``solver.step_src``)."""
    import ctypes
    from .. import _lib
    from ..solver import _ptr, _stream
    _lib.call("clawb200_sphere_src2", ctypes.byref(solver._problem), _ptr(state._q.cur), _ptr(state._aux.cur),
              float(dt), _stream())


def src2_torch(solver, state, dt):
    """The same source term written with tensor operations (what a user hook looks like).
    The radial unit vector that src2.f recomputes with mapc2p at every call is aux(14:16),
    bit for bit the same numbers."""
    q, aux = state.q, state.aux
    er0, er1, er2 = aux[13], aux[14], aux[15]
    six = torch.full((), 6.0, dtype=q.dtype, device=q.device)

    def project():
        qn = er0 * q[1] + er1 * q[2] + er2 * q[3]
        q[1] = q[1] - qn * er0
        q[2] = q[2] - qn * er1
        q[3] = q[3] - qn * er2
    project()
    fcor = _DF * er2
    RK = []
    hu, hv, hw = q[1], q[2], q[3]
    for st in range(4):
        if st > 0:
            hu = q[1] + 0.5 * RK[st - 1][0]
            hv = q[2] + 0.5 * RK[st - 1][1]
            hw = q[3] + 0.5 * RK[st - 1][2]
        RK.append((fcor * dt * (er2 * hv - er1 * hw),
                   dt * fcor * (er0 * hw - er2 * hu),
                   dt * fcor * (er1 * hu - er0 * hv)))
    for m in range(3):
        q[m + 1] = q[m + 1] + torch.div(RK[0][m] + 2.0 * RK[1][m] + 2.0 * RK[2][m] + RK[3][m], six)
    project()


def qbc_lower_y(state, dim, t, qbc, mbc):
    """pole fold: ghost row j mirrors interior row 2*mbc-1-j reversed in x
    (shallow_4_Rossby_Haurwitz_wave.py:292-300); rank-local with y-slabs"""
    for j in range(mbc):
        qbc[:, :, j] = torch.flip(qbc[:, :, 2 * mbc - 1 - j], dims=[1])


def qbc_upper_y(state, dim, t, qbc, mbc):
    my = state.grid.ng[1]
    for j in range(mbc):
        qbc[:, :, my + mbc + j] = torch.flip(qbc[:, :, my + mbc - 1 - j], dims=[1])


def setup(pyclaw, mx=40, my=20, mbc=2):
    """State + solver of the reference script (shallow_4_Rossby_Haurwitz_wave.py:330-440)."""
    xlower, xupper, ylower, yupper = -3.0, 1.0, -1.0, 1.0
    x = pyclaw.Dimension('x', xlower, xupper, mx)
    y = pyclaw.Dimension('y', ylower, yupper, my)
    grid = pyclaw.Grid([x, y])
    state = pyclaw.State(grid, 4, 16)
    dx, dy = grid.d
    j0, j1 = grid.y.nstart, grid.y.nend                  # this rank's slab
    ylo = ylower + j0 * dy
    auxfull = setaux(mbc, mx, j1 - j0, xlower, ylo, dx, dy)
    state.aux[:, :, :] = auxfull[:, mbc:-mbc, mbc:-mbc]
    state.q[:, :, :] = qinit(mx, j1 - j0, xlower, ylo, dx, dy)
    state.mcapa = 0
    state.aux_global['g'] = G_SW
    auxdev = torch.as_tensor(np.ascontiguousarray(auxfull.transpose(0, 2, 1)), device=state.device).permute(0, 2, 1)

    def auxbc_lower_y(state, dim, t, auxbc, mbc):
        auxbc[:, :, :mbc] = auxdev[:, :, :mbc]

    def auxbc_upper_y(state, dim, t, auxbc, mbc):
        auxbc[:, :, -mbc:] = auxdev[:, :, -mbc:]

    solver = pyclaw.ClawSolver2D()
    solver.rp = pyclaw.riemann.shallow_sphere
    solver.bc_lower[0] = solver.bc_upper[0] = pyclaw.BC.periodic
    solver.bc_lower[1] = solver.bc_upper[1] = pyclaw.BC.custom
    solver.user_bc_lower, solver.user_bc_upper = qbc_lower_y, qbc_upper_y
    solver.aux_bc_lower[0] = solver.aux_bc_upper[0] = pyclaw.BC.periodic
    solver.aux_bc_lower[1] = solver.aux_bc_upper[1] = pyclaw.BC.custom
    solver.user_aux_bc_lower, solver.user_aux_bc_upper = auxbc_lower_y, auxbc_upper_y
    solver.dim_split = 0
    solver.order_trans = 2
    solver.mwaves = 3
    solver.src_split = 2
    solver.step_src = src2
    solver.limiters = pyclaw.limiters.tvd.MC
    return state, solver
