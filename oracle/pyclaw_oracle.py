"""
CPU oracle driver -- TEST INFRASTRUCTURE ONLY (never imported by pyclaw_b200/).

numpy restatement of the Python-level control flow that surrounds the reference's
Fortran kernels, calling oracle/liboracle.so (claw_oracle.c) where the reference
calls its f2py modules.  Arrays are Fortran-ordered ``q[m,i,j]`` like the reference.

Follows (all under /root/reference/src/pyclaw):
  solver.py:315-452   apply_q_bcs / qbc_lower / qbc_upper  -> fill_bcs
  solver.py:602-717   evolve_to_time                        -> OracleSolver.evolve_to_time
  clawpack.py:114-165 ClawSolver.step                       -> OracleSolver._step_classic
  clawpack.py:299-324, 510-555 step_hyperbolic (1-D / 2-D)
  sharpclaw.py:152-237 SSP33 / SSP104 / Euler stages        -> OracleSolver._step_sharpclaw
  controller.py:195-303 Controller.run (outstyle 1)         -> run
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RP_ACOUSTICS, RP_ADVECTION, RP_EULER5, RP_SHALLOW, RP_SPHERE = 1, 2, 3, 4, 5
RP_NEL_FWAVE, RP_PSYSTEM = 6, 7   # f-wave solvers (step1fw.f / flux2fw.f corrections)
RP_ACOUSTICS3D_VC = 8              # 3-D variable-coefficient acoustics (dimensional splitting)
RP_VC_ACOUSTICS, RP_BURGERS, RP_ADVECTION_COLOR, RP_VC_ADVECTION, RP_EULER1D = 9, 10, 11, 12, 13
WENO_PYWENO_F32, WENO_PYWENO_F64, WENO_OLD, WENO_TABLES, RECON_TVD2 = 0, 1, 2, 3, 4
RECON_WENO_WAVE, RECON_WENO_FWAVE = 5, 6
BC_CUSTOM, BC_OUTFLOW, BC_PERIODIC, BC_REFLECTING = 0, 1, 2, 3

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build():
    """Compile oracle/liboracle.so with the committed Makefile."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        d, i = ctypes.c_double, ctypes.c_int
        L.oracle_step1.restype = d
        L.oracle_step1.argtypes = [i, _dp, i, i, i, i, i, _dp, _dp, d, d, _ip, _ip]
        L.oracle_step2ds.restype = d
        L.oracle_step2ds.argtypes = [i, _dp, i, i, i, i, i, i, i, _dp, _dp, _dp, d, d, d, _ip, _ip, i]
        L.oracle_step2.restype = d
        L.oracle_step2.argtypes = [i, _dp, i, i, i, i, i, i, i, _dp, _dp, _dp, d, d, d, _ip, _ip]
        L.oracle_set_weno_tables.restype = None
        L.oracle_set_weno_tables.argtypes = [i, _dp, _dp, _dp, _dp, _dp, d]
        L.oracle_step3ds.restype = d
        L.oracle_step3ds.argtypes = [i, _dp, i, i, i, i, i, i, i, _dp, _dp, _dp, d, d, d, d, _ip, _ip, i]
        L.oracle_step3.restype = d
        L.oracle_step3.argtypes = [i, _dp, i, i, i, i, i, i, i, _dp, _dp, _dp, d, d, d, d, _ip, _ip]
        L.oracle_sc_flux1.restype = d
        L.oracle_sc_flux1.argtypes = [i, _dp, i, i, i, i, _dp, _dp, d, d, i]
        L.oracle_sc_flux2.restype = d
        L.oracle_sc_flux2.argtypes = [i, _dp, i, i, i, i, i, _dp, _dp, d, d, d, i]
        L.oracle_sc_flux1_capa.restype = d
        L.oracle_sc_flux1_capa.argtypes = [i, _dp, i, i, i, i, _dp, _dp, d, d, i, _dp, i, i]
        L.oracle_sc_flux2_capa.restype = d
        L.oracle_sc_flux2_capa.argtypes = [i, _dp, i, i, i, i, i, _dp, _dp, d, d, d, i, _dp, i, i]
        L.oracle_step2_slabs.restype = d
        L.oracle_step2_slabs.argtypes = [i, _dp, i, i, i, i, i, i, _dp, _dp, _dp, d, d, d, _ip, _ip, i, i]
        L.oracle_sc_flux2_slabs.restype = d
        L.oracle_sc_flux2_slabs.argtypes = [i, _dp, i, i, i, i, i, _dp, _dp, d, d, d, i, i]
        L.oracle_sphere_setaux.restype = None
        L.oracle_sphere_setaux.argtypes = [i, i, i, d, d, d, d, _dp, d]
        L.oracle_sphere_qinit.restype = None
        L.oracle_sphere_qinit.argtypes = [i, i, i, d, d, d, d, _dp, d]
        L.oracle_sphere_src2.restype = None
        L.oracle_sphere_src2.argtypes = [i, i, d, d, d, d, _dp, _dp, d, d]
        L.oracle_set_tvd_limiters.restype = None
        L.oracle_set_tvd_limiters.argtypes = [_ip, i]
        L.oracle_rp_point.restype = None
        L.oracle_rp_point.argtypes = [i, _dp, i, i, i, ctypes.c_longlong, _dp, _dp, _dp, _dp, _dp, _dp,
                                      i, _dp, _dp, _dp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(_dp)


def _pi(a):
    return a.ctypes.data_as(_ip)


def _params(p):
    out = np.zeros(8)
    out[: len(p)] = p
    return out


# ---------------------------------------------------------------------------
# kernel-level wrappers (same array conventions as the f2py modules)
# ---------------------------------------------------------------------------
def step1(rp_id, rp_params, mbc, mx, qbc, auxbc, dx, dt, method, mthlim):
    """classic1.step1 (clawpack.py:323): updates qbc in place, returns cfl."""
    meqn = qbc.shape[0]
    maux = 0 if auxbc is None else auxbc.shape[0]
    aux = np.zeros(1) if maux == 0 else auxbc
    method = np.ascontiguousarray(method, dtype=np.int32)
    mthlim = np.ascontiguousarray(mthlim, dtype=np.int32)
    assert qbc.flags["F_CONTIGUOUS"]
    return lib().oracle_step1(rp_id, _p(_params(rp_params)), meqn, len(mthlim), mbc, maux, mx,
                              _p(qbc), _p(aux), dx, dt, _pi(method), _pi(mthlim))


def rp_point(rp_id, rp_params, ixy, mwaves, ql, qr, imp=0, asdq=None):
    """The restated Riemann solvers on n independent interfaces: ql, qr are [meqn, n] arrays
    (left / right state).  Returns (wave[meqn, mwaves, n], s[mwaves, n], amdq, apdq), or
    (bmasdq, bpasdq) of the transverse solver when ``asdq`` is given.  ixy = 0: 1-D solver."""
    ql = np.ascontiguousarray(ql, dtype=np.float64)
    qr = np.ascontiguousarray(qr, dtype=np.float64)
    meqn, n = ql.shape
    null = ctypes.cast(None, _dp)
    if asdq is None:
        wave = np.zeros((meqn, mwaves, n))
        s = np.zeros((mwaves, n))
        amdq, apdq = np.zeros((meqn, n)), np.zeros((meqn, n))
        lib().oracle_rp_point(rp_id, _p(_params(rp_params)), ixy, meqn, mwaves, n, _p(ql), _p(qr), _p(wave), _p(s),
                              _p(amdq), _p(apdq), 0, null, null, null)
        return wave, s, amdq, apdq
    asdq = np.ascontiguousarray(asdq, dtype=np.float64)
    bm, bp = np.zeros((meqn, n)), np.zeros((meqn, n))
    lib().oracle_rp_point(rp_id, _p(_params(rp_params)), ixy, meqn, mwaves, n, _p(ql), _p(qr), null, null, null,
                          null, imp, _p(asdq), _p(bm), _p(bp))
    return bm, bp


def step2ds(rp_id, rp_params, mbc, mx, my, qold, qnew, auxbc, dx, dy, dt, method, mthlim, ids):
    meqn = qold.shape[0]
    maux = 0 if auxbc is None else auxbc.shape[0]
    aux = np.zeros(1) if maux == 0 else auxbc
    method = np.ascontiguousarray(method, dtype=np.int32)
    mthlim = np.ascontiguousarray(mthlim, dtype=np.int32)
    assert qold.flags["F_CONTIGUOUS"] and qnew.flags["F_CONTIGUOUS"]
    return lib().oracle_step2ds(rp_id, _p(_params(rp_params)), max(mx, my), meqn, len(mthlim), maux,
                                mbc, mx, my, _p(qold), _p(qnew), _p(aux), dx, dy, dt,
                                _pi(method), _pi(mthlim), ids)


def step2(rp_id, rp_params, mbc, mx, my, qold, qnew, auxbc, dx, dy, dt, method, mthlim):
    meqn = qold.shape[0]
    maux = 0 if auxbc is None else auxbc.shape[0]
    aux = np.zeros(1) if maux == 0 else auxbc
    method = np.ascontiguousarray(method, dtype=np.int32)
    mthlim = np.ascontiguousarray(mthlim, dtype=np.int32)
    assert qold.flags["F_CONTIGUOUS"] and qnew.flags["F_CONTIGUOUS"]
    return lib().oracle_step2(rp_id, _p(_params(rp_params)), max(mx, my), meqn, len(mthlim), maux,
                              mbc, mx, my, _p(qold), _p(qnew), _p(aux), dx, dy, dt,
                              _pi(method), _pi(mthlim))


def step3ds(rp_id, rp_params, mbc, mx, my, mz, qold, qnew, auxbc, dx, dy, dz, dt, method, mthlim, idir):
    """classic3.step3ds (clawpack.py:656-676): one directional sweep; qnew updated in place."""
    meqn = qold.shape[0]
    maux = 0 if auxbc is None else auxbc.shape[0]
    aux = np.zeros(1) if maux == 0 else auxbc
    method = np.ascontiguousarray(method, dtype=np.int32)
    mthlim = np.ascontiguousarray(mthlim, dtype=np.int32)
    assert qold.flags["F_CONTIGUOUS"] and qnew.flags["F_CONTIGUOUS"]
    return lib().oracle_step3ds(rp_id, _p(_params(rp_params)), meqn, len(mthlim), maux, mbc, mx, my, mz,
                                _p(qold), _p(qnew), _p(aux), dx, dy, dz, dt, _pi(method), _pi(mthlim), idir)


def step3(rp_id, rp_params, mbc, mx, my, mz, qold, qnew, auxbc, dx, dy, dz, dt, method, mthlim):
    """classic3.step3 (clawpack.py:680-682; step3.f + flux3.f): the unsplit 3-D step; qnew (== qold
    on entry) is updated in place.  method[2] = 0 | 10 | 11 | 20 | 21 | 22."""
    meqn = qold.shape[0]
    maux = 0 if auxbc is None else auxbc.shape[0]
    aux = np.zeros(1) if maux == 0 else auxbc
    method = np.ascontiguousarray(method, dtype=np.int32)
    mthlim = np.ascontiguousarray(mthlim, dtype=np.int32)
    assert qold.flags["F_CONTIGUOUS"] and qnew.flags["F_CONTIGUOUS"]
    cfl = lib().oracle_step3(rp_id, _p(_params(rp_params)), meqn, len(mthlim), maux, mbc, mx, my, mz,
                             _p(qold), _p(qnew), _p(aux), dx, dy, dz, dt, _pi(method), _pi(mthlim))
    if cfl < 0:
        raise ValueError("oracle_step3: unsupported solver / capa")
    return cfl


def step2_slabs(rp_id, rp_params, mbc, mx, my, qold, qnew, auxbc, dx, dy, dt, method, mthlim,
                nthreads, dimsplit):
    """Host-parallel (y-slab) classic step; identical results to step2/step2ds."""
    meqn = qold.shape[0]
    maux = 0 if auxbc is None else auxbc.shape[0]
    aux = np.zeros(1) if maux == 0 else auxbc
    method = np.ascontiguousarray(method, dtype=np.int32)
    mthlim = np.ascontiguousarray(mthlim, dtype=np.int32)
    return lib().oracle_step2_slabs(rp_id, _p(_params(rp_params)), meqn, len(mthlim), maux, mbc,
                                    mx, my, _p(qold), _p(qnew), _p(aux), dx, dy, dt,
                                    _pi(method), _pi(mthlim), nthreads, int(dimsplit))


def set_tvd_limiters(mthlim):
    """Limiter per COMPONENT for weno_variant = RECON_TVD2 (reconstruct.f90:568-625)."""
    a = np.ascontiguousarray(mthlim, dtype=np.int32)
    lib().oracle_set_tvd_limiters(_pi(a), len(a))


def set_weno_tables(tab):
    """Coefficient tables for weno_variant = WENO_TABLES (orders 7..17): dict with k, S[k][k(k+1)/2],
    CL/CR[k][k], WL/WR[k], eps (see oracle_set_weno_tables in claw_oracle.c)."""
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    S, CL, CR, WL, WR = c(tab['S']), c(tab['CL']), c(tab['CR']), c(tab['WL']), c(tab['WR'])
    lib().oracle_set_weno_tables(int(tab['k']), _p(S), _p(CL), _p(CR), _p(WL), _p(WR), float(tab['eps']))


def sc_flux1(rp_id, rp_params, mwaves, mbc, mx, q, dx, dt, weno_variant, auxbc=None, mcapa=0):
    """sharpclaw1.flux1 (sharpclaw.py:385); mcapa is 1-based as in clawparams.mcapa (0 = none)."""
    dq = np.zeros_like(q, order="F")
    if auxbc is not None:      # capacity function and / or a solver that reads aux
        cfl = lib().oracle_sc_flux1_capa(rp_id, _p(_params(rp_params)), q.shape[0], mwaves, mbc, mx,
                                         _p(q), _p(dq), dx, dt, weno_variant, _p(auxbc), auxbc.shape[0], mcapa)
        return dq, cfl
    cfl = lib().oracle_sc_flux1(rp_id, _p(_params(rp_params)), q.shape[0], mwaves, mbc, mx,
                                _p(q), _p(dq), dx, dt, weno_variant)
    return dq, cfl


def sc_flux2(rp_id, rp_params, mwaves, mbc, mx, my, q, dx, dy, dt, weno_variant, nthreads=1,
             auxbc=None, mcapa=0):
    dq = np.zeros_like(q, order="F")
    if auxbc is not None:
        cfl = lib().oracle_sc_flux2_capa(rp_id, _p(_params(rp_params)), q.shape[0], mwaves, mbc, mx, my,
                                         _p(q), _p(dq), dx, dy, dt, weno_variant, _p(auxbc),
                                         auxbc.shape[0], mcapa)
        return dq, cfl
    if nthreads > 1:
        cfl = lib().oracle_sc_flux2_slabs(rp_id, _p(_params(rp_params)), q.shape[0], mwaves, mbc,
                                          mx, my, _p(q), _p(dq), dx, dy, dt, weno_variant, nthreads)
    else:
        cfl = lib().oracle_sc_flux2(rp_id, _p(_params(rp_params)), q.shape[0], mwaves, mbc, mx, my,
                                    _p(q), _p(dq), dx, dy, dt, weno_variant)
    return dq, cfl


# ---------------------------------------------------------------------------
# shallow water on the sphere: application helpers (apps/shallow-sphere/*.f)
# ---------------------------------------------------------------------------
def sphere_setaux(mbc, mx, my, xlower, ylower, dx, dy, Rsphere=1.0):
    aux = np.zeros((16, mx + 2 * mbc, my + 2 * mbc), order="F")
    lib().oracle_sphere_setaux(mbc, mx, my, xlower, ylower, dx, dy, _p(aux), Rsphere)
    return aux


def sphere_qinit(mbc, mx, my, xlower, ylower, dx, dy, Rsphere=1.0):
    q = np.zeros((4, mx + 2 * mbc, my + 2 * mbc), order="F")
    lib().oracle_sphere_qinit(mbc, mx, my, xlower, ylower, dx, dy, _p(q), Rsphere)
    return q


def sphere_src2(q, aux, xlower, ylower, dx, dy, dt, Rsphere=1.0):
    """src2.f on interior arrays; q is updated in place (must be F-contiguous)."""
    assert q.flags["F_CONTIGUOUS"] and aux.flags["F_CONTIGUOUS"]
    lib().oracle_sphere_src2(q.shape[1], q.shape[2], xlower, ylower, dx, dy, _p(q), _p(aux), dt, Rsphere)


# ---------------------------------------------------------------------------
# boundary conditions (solver.py:354-452)
# ---------------------------------------------------------------------------
def fill_bcs(qbc, mbc, bc_lower, bc_upper, user_lower=None, user_upper=None, t=0.0, negate=True):
    ndim = qbc.ndim - 1
    for idim in range(ndim):
        v = np.rollaxis(qbc, idim + 1, 1)
        b = bc_lower[idim]
        if b == BC_CUSTOM:
            user_lower(idim, t, qbc, mbc)
        elif b == BC_OUTFLOW:
            for i in range(mbc):
                v[:, i, ...] = v[:, mbc, ...]
        elif b == BC_PERIODIC:
            v[:, :mbc, ...] = v[:, -2 * mbc:-mbc, ...]
        elif b == BC_REFLECTING:
            for i in range(mbc):
                v[:, i, ...] = v[:, 2 * mbc - 1 - i, ...]
                if negate:
                    v[idim + 1, i, ...] = -v[idim + 1, 2 * mbc - 1 - i, ...]
        else:
            raise NotImplementedError(b)
        b = bc_upper[idim]
        if b == BC_CUSTOM:
            user_upper(idim, t, qbc, mbc)
        elif b == BC_OUTFLOW:
            for i in range(mbc):
                v[:, -i - 1, ...] = v[:, -mbc - 1, ...]
        elif b == BC_PERIODIC:
            v[:, -mbc:, ...] = v[:, mbc:2 * mbc, ...]
        elif b == BC_REFLECTING:
            for i in range(mbc):
                v[:, -i - 1, ...] = v[:, -2 * mbc + i, ...]
                if negate:
                    v[idim + 1, -i - 1, ...] = -v[idim + 1, -2 * mbc + i, ...]
        else:
            raise NotImplementedError(b)


def _interior(a, mbc):
    return a[(slice(None),) + (slice(mbc, -mbc),) * (a.ndim - 1)]


class OracleSolver(object):
    """One object for the four solver classes; ``kind`` is 'classic' or 'sharpclaw'."""

    def __init__(self, kind, ndim, rp_id, rp_params, mwaves):
        self.kind, self.ndim, self.rp_id, self.rp_params, self.mwaves = kind, ndim, rp_id, rp_params, mwaves
        self.dt_initial, self.dt_max, self.max_steps, self.dt_variable = 0.1, 1e99, 1000, True
        self.bc_lower, self.bc_upper = [None] * ndim, [None] * ndim
        self.aux_bc_lower, self.aux_bc_upper = [BC_OUTFLOW] * ndim, [BC_OUTFLOW] * ndim
        self.user_bc_lower = self.user_bc_upper = None
        self.user_aux_bc_lower = self.user_aux_bc_upper = None
        self.step_src = None
        self.src_split = 1
        if kind == "classic":
            self.mbc, self.cfl_max, self.cfl_desired = 2, 1.0, 0.9
            self.limiters, self.order, self.dim_split, self.order_trans = 1, 2, True, 1
        else:
            self.mbc, self.cfl_max, self.cfl_desired = 3, 2.5, 2.45
            self.time_integrator, self.weno_variant = "SSP104", WENO_PYWENO_F32
            self.weno_order, self.weno_tables = 5, None
            self.lim_type, self.limiters = 2, [1]
            self.char_decomp, self.fwave = 0, False
        self.mcapa = -1
        self.cfl = self.cfl_desired
        self.status = {}
        self.nthreads = 1

    # ---- setup (clawpack.py:214-238, sharpclaw.py:303-326) ----
    def setup(self, q, aux, d):
        self.d = list(d)
        if self.kind == "sharpclaw" and self.weno_order != 5:
            # sharpclaw.py:303-304 mbc = (weno_order+1)/2 ; reconstruct.f90:96-113 selects weno<order>;
            # the caller supplies the coefficient tables (set_weno_tables)
            self.mbc = (self.weno_order + 1) // 2
            set_weno_tables(self.weno_tables)
            self.weno_variant = WENO_TABLES
        if self.kind == "sharpclaw" and self.lim_type == 2 and self.char_decomp == 1:
            assert self.ndim == 1  # the reference's 2-D flux1 cannot run this branch (wrong rpn2 argument list)
            self.weno_variant = RECON_WENO_FWAVE if self.fwave else RECON_WENO_WAVE
        if self.kind == "sharpclaw" and self.lim_type == 1:
            # flux1.f90:79-83 tvd2; clawparams.mthlim = solver.mthlim (sharpclaw.py:213-218, 278)
            lim = self.limiters if isinstance(self.limiters, list) else [self.limiters]
            if len(lim) == 1:
                lim = lim * self.mwaves
            set_tvd_limiters(lim)
            self.weno_variant = RECON_TVD2
        mbc = self.mbc
        self.n = list(q.shape[1:])
        self.qbc = np.zeros([q.shape[0]] + [n + 2 * mbc for n in self.n], order="F")
        if aux is not None:
            self.auxbc = np.zeros([aux.shape[0]] + [n + 2 * mbc for n in self.n], order="F")
            _interior(self.auxbc, mbc)[...] = aux
            fill_bcs(self.auxbc, mbc, self.aux_bc_lower, self.aux_bc_upper, self.user_aux_bc_lower,
                     self.user_aux_bc_upper, negate=False)
        else:
            self.auxbc = None
        if self.kind == "classic":
            lim = self.limiters
            if not isinstance(lim, list):
                lim = [lim]
            if len(lim) == 1:
                lim = lim * self.mwaves
            self.mthlim = np.array(lim, dtype=np.int32)
            trans = 0 if self.ndim == 1 else (-1 if self.dim_split else self.order_trans)
            self.method = np.array([int(self.dt_variable), self.order, trans, 0, 0, self.mcapa + 1,
                                    0 if aux is None else aux.shape[0]], dtype=np.int32)
        self.dt = self.dt_initial
        self.cfl = self.cfl_desired

    def _bcs(self, q, t):
        _interior(self.qbc, self.mbc)[...] = q
        fill_bcs(self.qbc, self.mbc, self.bc_lower, self.bc_upper, self.user_bc_lower,
                 self.user_bc_upper, t)

    # ---- classic ----
    def _hyperbolic_classic(self, state):
        self._bcs(state["q"], state["t"])
        mbc = self.mbc
        if self.ndim == 1:
            cfl = step1(self.rp_id, self.rp_params, mbc, self.n[0], self.qbc, self.auxbc,
                        self.d[0], self.dt, self.method, self.mthlim)
        elif self.ndim == 3:
            # clawpack.py:656-682: three aliased step3ds calls, or one step3 call
            mx, my, mz = self.n
            dx, dy, dz = self.d
            q = self.qbc
            cfl = 0.0
            if not self.dim_split:
                qold = q.copy("F")
                cfl = step3(self.rp_id, self.rp_params, mbc, mx, my, mz, qold, q, self.auxbc,
                            dx, dy, dz, self.dt, self.method, self.mthlim)
            for idir in ((1, 2, 3) if self.dim_split else ()):
                qold = q.copy("F")
                cfl = max(cfl, step3ds(self.rp_id, self.rp_params, mbc, mx, my, mz, qold, q, self.auxbc,
                                       dx, dy, dz, self.dt, self.method, self.mthlim, idir))
        else:
            mx, my = self.n
            dx, dy = self.d
            qnew = self.qbc
            qold = qnew.copy("F")
            if self.nthreads > 1:
                cfl = step2_slabs(self.rp_id, self.rp_params, mbc, mx, my, qold, qnew, self.auxbc,
                                  dx, dy, self.dt, self.method, self.mthlim, self.nthreads,
                                  self.dim_split)
            elif self.dim_split:
                cx = step2ds(self.rp_id, self.rp_params, mbc, mx, my, qold, qnew, self.auxbc,
                             dx, dy, self.dt, self.method, self.mthlim, 1)
                cy = step2ds(self.rp_id, self.rp_params, mbc, mx, my, qnew, qnew, self.auxbc,
                             dx, dy, self.dt, self.method, self.mthlim, 2)
                cfl = max(cx, cy)
            else:
                cfl = step2(self.rp_id, self.rp_params, mbc, mx, my, qold, qnew, self.auxbc,
                            dx, dy, self.dt, self.method, self.mthlim)
        self.cfl = cfl
        state["q"] = _interior(self.qbc, mbc).copy("F")

    def _step_classic(self, state):
        if self.src_split == 2 and self.step_src is not None:
            self.step_src(self, state, self.dt / 2.0)
        self._hyperbolic_classic(state)
        if self.cfl >= self.cfl_max:
            return False
        if self.step_src is not None:
            if self.src_split == 2:
                self.step_src(self, state, self.dt / 2.0)
            if self.src_split == 1:
                self.step_src(self, state, self.dt)
        return True

    # ---- sharpclaw ----
    def _dq(self, q, t):
        self._bcs(q, t)
        mbc = self.mbc
        if self.ndim == 1:
            dq, cfl = sc_flux1(self.rp_id, self.rp_params, self.mwaves, mbc, self.n[0], self.qbc,
                               self.d[0], self.dt, self.weno_variant, auxbc=self.auxbc,
                               mcapa=self.mcapa + 1)
        else:
            dq, cfl = sc_flux2(self.rp_id, self.rp_params, self.mwaves, mbc, self.n[0], self.n[1],
                               self.qbc, self.d[0], self.d[1], self.dt, self.weno_variant,
                               self.nthreads, auxbc=self.auxbc, mcapa=self.mcapa + 1)
        self.cfl = cfl
        if cfl > self.cfl_max:
            raise _CFLError()
        return _interior(dq, mbc)

    def _step_sharpclaw(self, state):
        q, t = state["q"], state["t"]
        try:
            if self.time_integrator == "Euler":
                state["q"] = q + self._dq(q, t)
            elif self.time_integrator == "SSP33":
                s = q + self._dq(q, t)
                s = 0.75 * q + 0.25 * (s + self._dq(s, t))
                state["q"] = 1. / 3. * q + 2. / 3. * (s + self._dq(s, t))
            elif self.time_integrator == "SSP104":
                s1 = q + self._dq(q, t) / 6.
                for _ in range(4):
                    s1 = s1 + self._dq(s1, t) / 6.
                s2 = q / 25. + 9. / 25 * s1
                s1 = 15. * s2 - 5. * s1
                for _ in range(4):
                    s1 = s1 + self._dq(s1, t) / 6.
                state["q"] = s2 + 0.6 * s1 + 0.1 * self._dq(s1, t)
            else:
                raise ValueError(self.time_integrator)
        except _CFLError:
            return False

    def step(self, state):
        if self.kind == "classic":
            return self._step_classic(state)
        return self._step_sharpclaw(state)

    # ---- solver.py:602-717 ----
    def evolve_to_time(self, state, tend):
        tstart = state["t"]
        self.status = {"cflmax": self.cfl, "dtmin": self.dt, "dtmax": self.dt, "numsteps": 0,
                       "rejected": 0}
        max_steps = self.max_steps
        if not self.dt_variable:
            max_steps = int((tend - tstart + 1e-10) / self.dt)
        if tend <= tstart:
            max_steps = 0
        for n in range(max_steps):
            if state["t"] + self.dt > tend and tstart < tend:
                self.dt = tend - state["t"]
            if self.dt_variable:
                q_backup = state["q"].copy("F")
                told = state["t"]
            self.step(state)
            cfl = self.cfl
            if cfl <= self.cfl_max:
                self.status["cflmax"] = max(cfl, self.status["cflmax"])
                if self.dt_variable:
                    state["t"] += self.dt
                else:
                    state["t"] = tstart + (n + 1) * self.dt
                self.status["numsteps"] += 1
                if state["t"] >= tend:
                    break
            else:
                if self.dt_variable:
                    state["q"] = q_backup
                    state["t"] = told
                    self.status["rejected"] += 1
                else:
                    raise Exception("CFL too large, giving up!")
            if self.dt_variable:
                if cfl > 0.0:
                    self.dt = min(self.dt_max, self.dt * self.cfl_desired / cfl)
                    self.status["dtmin"] = min(self.dt, self.status["dtmin"])
                    self.status["dtmax"] = max(self.dt, self.status["dtmax"])
                else:
                    self.dt = self.dt_max
        return self.status

    # ---- controller.py:195-303, outstyle 1 ----
    def run(self, q0, aux, d, tfinal, nout, t0=0.0):
        state = {"q": np.array(q0, order="F", copy=True), "aux": aux, "t": t0}
        self.dt = self.dt_initial
        self.setup(state["q"], aux, d)
        frames = [state["q"].copy("F")]
        total = {"numsteps": 0, "rejected": 0}
        for t in np.linspace(t0, tfinal, nout + 1)[1:]:
            st = self.evolve_to_time(state, t)
            total["numsteps"] += st["numsteps"]
            total["rejected"] += st["rejected"]
            frames.append(state["q"].copy("F"))
        self.total = total
        return frames


class _CFLError(Exception):
    pass
