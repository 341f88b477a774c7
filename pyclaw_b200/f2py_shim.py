"""Stand-ins for the reference's f2py extension modules, over the host-pointer C ABI.

``classic1``, ``classic2``, ``sharpclaw1``, ``sharpclaw2`` objects with the call signatures the
reference's clawpack.py / sharpclaw.py use (path B).  The Riemann solver,
which the reference fixes at link time (RP_SOURCE in the app Makefile), is chosen when the
module object is created::

    classic2 = f2py_shim.classic2('euler_5wave')
    classic2.cparam.gamma, classic2.cparam.gamma1 = 1.4, 0.4
    qnew, cfl = classic2.step2(maxm, mbc, mx, my, qold, qnew, auxbc, dx, dy, dt, method, mthlim,
                               aux1, aux2, aux3, work)
"""
import ctypes

import numpy as np

from . import _lib, riemann


class _CParam(object):
    """the `cparam` common block, filled by State.set_cparam (state.py:142-162)"""
    pass


class _Module(object):
    def __init__(self, rp, ndim):
        self._rp = riemann.resolve(rp, {}, ndim)
        self._ndim = ndim
        self.cparam = _CParam()
        self.weno_variant = _lib.WENO_PYWENO_F32

    def _problem(self, mbc, mx, my, q, dx, dy, method, mthlim, auxbc=None):
        params = [float(getattr(self.cparam, k, self._rp.defaults.get(k, 0.0))) for k in self._rp.param_names]
        maux = 0 if self._aux(auxbc) is None else auxbc.shape[0]
        if method is None and maux:
            method = [1, 2, 0, 0, 0, int(getattr(self, 'mcapa', 0)), maux]   # clawparams.mcapa (1-based)
        return _lib.make_problem(self._ndim, q.shape[0], self._rp.nwaves(self._ndim), mbc, mx, my, dx, dy,
                                 self._rp.rp_id, params, method=method, mthlim=mthlim, maux=maux,
                                 weno_variant=self.weno_variant)

    def _aux(self, auxbc):
        """f2py passes a dummy array when the application has no aux; None here."""
        if auxbc is None or not isinstance(auxbc, np.ndarray) or auxbc.ndim != self._ndim + 1 or auxbc.shape[0] == 0:
            return None
        return self._check(auxbc)

    @staticmethod
    def _check(a):
        if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags['F_CONTIGUOUS']):
            raise ValueError("expected a Fortran-ordered float64 array (what f2py would require)")
        return ctypes.c_void_p(a.ctypes.data)


class classic1(_Module):
    def __init__(self, rp):
        super().__init__(rp, 1)

    def step1(self, mbc, mx, qbc, auxbc, dx, dt, method, mthlim):
        P, cfl = self._problem(mbc, mx, 1, qbc, dx, 1.0, list(method), list(mthlim), auxbc), ctypes.c_double()
        _lib.call("clawb200_step1_host", ctypes.byref(P), self._check(qbc), self._aux(auxbc), float(dt),
                  ctypes.byref(cfl))
        return qbc, cfl.value


class classic2(_Module):
    def __init__(self, rp):
        super().__init__(rp, 2)

    def step2ds(self, maxm, mbc, mx, my, qold, qnew, auxbc, dx, dy, dt, method, mthlim,
                aux1=None, aux2=None, aux3=None, work=None, ids=1):
        P, cfl = self._problem(mbc, mx, my, qold, dx, dy, list(method), list(mthlim), auxbc), ctypes.c_double()
        _lib.call("clawb200_step2ds_host", ctypes.byref(P), self._check(qold), self._check(qnew), self._aux(auxbc),
                  float(dt), int(ids), ctypes.byref(cfl))
        return qnew, cfl.value

    def step2(self, maxm, mbc, mx, my, qold, qnew, auxbc, dx, dy, dt, method, mthlim,
              aux1=None, aux2=None, aux3=None, work=None):
        P, cfl = self._problem(mbc, mx, my, qold, dx, dy, list(method), list(mthlim), auxbc), ctypes.c_double()
        _lib.call("clawb200_step2_host", ctypes.byref(P), self._check(qold), self._check(qnew), self._aux(auxbc),
                  float(dt), ctypes.byref(cfl))
        return qnew, cfl.value


class sharpclaw1(_Module):
    def __init__(self, rp):
        super().__init__(rp, 1)

    def flux1(self, q, auxbc, dt, t, ixy, mx, mbc, maxnx, dx=None):
        dq = np.zeros_like(q, order='F')
        P, cfl = self._problem(mbc, mx, 1, q, dx, 1.0, None, None, auxbc), ctypes.c_double()
        _lib.call("clawb200_sharpclaw_dq_host", ctypes.byref(P), self._check(q), self._check(dq), self._aux(auxbc),
                  float(dt), ctypes.byref(cfl))
        return dq, cfl.value


class sharpclaw2(_Module):
    def __init__(self, rp):
        super().__init__(rp, 2)

    def flux2(self, q, auxbc, dt, t, mbc, maxm, mx, my, dx=None, dy=None):
        dq = np.zeros_like(q, order='F')
        P, cfl = self._problem(mbc, mx, my, q, dx, dy, None, None, auxbc), ctypes.c_double()
        _lib.call("clawb200_sharpclaw_dq_host", ctypes.byref(P), self._check(q), self._check(dq), self._aux(auxbc),
                  float(dt), ctypes.byref(cfl))
        return dq, cfl.value


class classic3(_Module):
    """classic3.step3ds (clawpack.py:656-676); only the dimensionally split routine exists."""

    def __init__(self, rp='vc_acoustics_3d'):
        super().__init__(rp, 3)

    def step3ds(self, maxm, mbc, mx, my, mz, qold, qnew, auxbc, dx, dy, dz, dt, method, mthlim,
                aux1=None, aux2=None, aux3=None, work=None, idir=1):
        P, cfl = self._problem(mbc, mx, my, qold, dx, dy, list(method), list(mthlim), auxbc), ctypes.c_double()
        # qold may be qnew (the reference's second and third calls pass the same array twice):
        # the library uploads qold before it writes qnew
        _lib.call("clawb200_step3ds_host", ctypes.byref(P), int(mz), float(dz), self._check(qold), self._check(qnew),
                  self._aux(auxbc), float(dt), int(idir), ctypes.byref(cfl))
        return qnew, cfl.value
