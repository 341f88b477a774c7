"""The CUDA Riemann solvers (the device functions inlined into the sweeps), called through
`clawb200_rp_solve_host` / `clawb200_rp_transverse_host`: (1) against oracle-independent
properties (tests/rp_properties.py), (2) bit for bit against the oracle's solvers."""
import ctypes

import numpy as np
import pytest

import rp_properties as rpp
import test_rp_properties as cpu
from oracle import pyclaw_oracle as po
from pyclaw_b200 import _lib

pytestmark = pytest.mark.gpu

MEQN = {po.RP_EULER5: 5, po.RP_SHALLOW: 3, po.RP_ACOUSTICS: 3, po.RP_ADVECTION: 1}


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _problem(name):
    rp_id, params, mw = cpu._spec(name)
    return _lib.make_problem(2, MEQN[rp_id], mw, 2, 8, 8, 1.0, 1.0, rp_id, params), mw


def _solve_for(name):
    P, mw = _problem(name)

    def solve(ixy, ql, qr):
        ql, qr = np.ascontiguousarray(ql), np.ascontiguousarray(qr)
        meqn, n = ql.shape
        wave, s = np.zeros((meqn, mw, n)), np.zeros((mw, n))
        amdq, apdq = np.zeros((meqn, n)), np.zeros((meqn, n))
        _lib.call("clawb200_rp_solve_host", ctypes.byref(P), ixy, n, _ptr(ql), _ptr(qr), None, None, _ptr(wave), _ptr(s),
                  _ptr(amdq), _ptr(apdq))
        return wave, s, amdq, apdq
    return solve


def _transverse_for(name):
    P, mw = _problem(name)

    def transverse(ixy, ql, qr, imp, asdq):
        ql, qr, asdq = np.ascontiguousarray(ql), np.ascontiguousarray(qr), np.ascontiguousarray(asdq)
        meqn, n = ql.shape
        bm, bp = np.zeros((meqn, n)), np.zeros((meqn, n))
        _lib.call("clawb200_rp_transverse_host", ctypes.byref(P), ixy, n, _ptr(ql), _ptr(qr), imp, _ptr(asdq),
                  _ptr(bm), _ptr(bp))
        return bm, bp
    return transverse


def test_cuda_riemann_solvers_satisfy_their_defining_properties():
    rpp.run_all(_solve_for, _transverse_for, n=8192)


@pytest.mark.parametrize("name", ["euler", "shallow", "acoustics", ("advection", (0.7, -0.4))])
@pytest.mark.parametrize("ixy", [1, 2])
def test_cuda_riemann_solvers_equal_the_oracle_bit_for_bit(name, ixy):
    n = 4096
    if name == "euler":
        ql, qr, (sl, unl, unr) = rpp.euler_states(n, 11)
        rpp._set_normal(ql, ixy, sl, unl, 1.0)   # make the transonic block physical (p > 0)
        rpp._set_normal(qr, ixy, sl, unr, 1.0)
    elif name == "shallow":
        ql, qr = rpp.shallow_states(n, 12)
    else:
        rng = np.random.RandomState(13)
        m = 3 if name == "acoustics" else 1
        ql, qr = rng.uniform(-1, 1, (m, n)), rng.uniform(-1, 1, (m, n))
    got = _solve_for(name)(ixy, ql, qr)
    want = cpu._solve_for(name)(ixy, ql, qr)
    for g, w, what in zip(got, want, ("wave", "s", "amdq", "apdq")):
        assert np.isfinite(w).all(), (name, what)
        assert np.array_equal(g, w), (name, ixy, what, np.abs(g - w).max())
    asdq = np.random.RandomState(14).uniform(-1, 1, ql.shape)
    for imp in (1, 2):
        gb = _transverse_for(name)(ixy, ql, qr, imp, asdq)
        wb = cpu._transverse_for(name)(ixy, ql, qr, imp, asdq)
        assert np.array_equal(gb[0], wb[0]) and np.array_equal(gb[1], wb[1]), (name, ixy, imp)


@pytest.mark.parametrize("law", [1, 2])
def test_fwave_elasticity_solver_properties(law):
    """The f-wave solver of the stegoton / p-system applications (external, no golden in the
    reference; its own psystem test is `return True`): the f-waves sum to the flux difference
    f(q_r; aux_r) - f(q_l; aux_l) with f = (-u, -sigma(eps)), the fluctuations are the f-waves,
    and each f-wave is parallel to the eigenvector (+-Z, 1) of its side's impedance."""
    n = 4096
    rng = np.random.RandomState(law)
    rho_l, rho_r = rng.uniform(1, 4, n), rng.uniform(1, 4, n)
    E_l, E_r = rng.uniform(1, 4, n), rng.uniform(1, 4, n)
    eps_l, eps_r = rng.uniform(-0.2, 0.5, n), rng.uniform(-0.2, 0.5, n)
    ql = np.stack([eps_l, rho_l * rng.uniform(-1, 1, n)])
    qr = np.stack([eps_r, rho_r * rng.uniform(-1, 1, n)])
    auxl = np.ascontiguousarray(np.stack([rho_l, E_l, np.zeros(n)]))
    auxr = np.ascontiguousarray(np.stack([rho_r, E_r, np.zeros(n)]))
    P = _lib.make_problem(1, 2, 2, 2, 8, 1, 1.0, 1.0, po.RP_NEL_FWAVE, [float(law)], maux=3)
    wave, s = np.zeros((2, 2, n)), np.zeros((2, n))
    amdq, apdq = np.zeros((2, n)), np.zeros((2, n))
    _lib.call("clawb200_rp_solve_host", ctypes.byref(P), 1, n, _ptr(np.ascontiguousarray(ql)),
              _ptr(np.ascontiguousarray(qr)), _ptr(auxl), _ptr(auxr), _ptr(wave), _ptr(s), _ptr(amdq), _ptr(apdq))
    sig = (lambda e, E: E * e) if law == 1 else (lambda e, E: np.exp(E * e) - 1.0)
    sigp = (lambda e, E: E + 0 * e) if law == 1 else (lambda e, E: E * np.exp(E * e))
    df = np.stack([-(qr[1] / rho_r - ql[1] / rho_l), -(sig(eps_r, E_r) - sig(eps_l, E_l))])
    assert np.abs(wave.sum(axis=1) - df).max() < 1e-13
    assert np.array_equal(amdq, wave[:, 0]) and np.array_equal(apdq, wave[:, 1])
    cl, cr = np.sqrt(sigp(eps_l, E_l) / rho_l), np.sqrt(sigp(eps_r, E_r) / rho_r)
    assert np.abs(s[0] + cl).max() < 1e-14 and np.abs(s[1] - cr).max() < 1e-14
    # left-going f-wave along (1, Z_l), right-going along (1, -Z_r)
    assert np.abs(wave[1, 0] - wave[0, 0] * cl * rho_l).max() < 1e-13
    assert np.abs(wave[1, 1] + wave[0, 1] * cr * rho_r).max() < 1e-13
