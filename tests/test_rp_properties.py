"""The oracle's Riemann solvers against oracle-independent properties (tests/rp_properties.py)."""
import numpy as np

import rp_properties as rpp
from oracle import pyclaw_oracle as po

SOLVERS = {"euler": (po.RP_EULER5, [rpp.GAMMA, rpp.GAMMA1], 5), "shallow": (po.RP_SHALLOW, [rpp.GRAV], 3),
           "acoustics": (po.RP_ACOUSTICS, [1.0, 4.0, 2.0, 2.0], 2)}


def _spec(name):
    if isinstance(name, tuple):
        return po.RP_ADVECTION, list(name[1]), 1
    return SOLVERS[name]


def _solve_for(name):
    rp_id, params, mw = _spec(name)
    return lambda ixy, ql, qr: po.rp_point(rp_id, params, ixy, mw, ql, qr)


def _transverse_for(name):
    rp_id, params, mw = _spec(name)
    return lambda ixy, ql, qr, imp, asdq: po.rp_point(rp_id, params, ixy, mw, ql, qr, imp, asdq)


def test_oracle_riemann_solvers_satisfy_their_defining_properties():
    rpp.run_all(_solve_for, _transverse_for, n=2048)


def test_property_checks_detect_a_wrong_solver():
    """The checks are not vacuous: a solver with one wave speed off by 1e-6, or a transverse
    solver that splits with the wrong sign convention, fails them."""
    import pytest
    ql, qr = rpp.shallow_states(256, 5)
    u, v, c2 = rpp.shallow_roe(ql, qr)
    jac = lambda d: rpp.shallow_jacobian(u, v, c2, d)
    good = _solve_for("shallow")

    def bad_speed(ixy, l, r):
        w, s, am, ap = good(ixy, l, r)
        s = s.copy()
        s[0] += 1e-6
        return w, s, am, ap
    with pytest.raises(AssertionError):
        rpp.check_normal("bad", bad_speed, 1, ql, qr, rpp.shallow_flux, jac)
    tg = _transverse_for("shallow")
    swapped = lambda ixy, l, r, imp, a: tg(ixy, l, r, imp, a)[::-1]
    c = np.sqrt(c2)
    eig = lambda d: [(u if d == 1 else v) - c, (u if d == 1 else v), (u if d == 1 else v) + c]
    asdq = np.random.RandomState(0).uniform(-1, 1, ql.shape)
    rpp.check_transverse("good", tg, 1, ql, qr, asdq, jac, eig)
    with pytest.raises(AssertionError):
        rpp.check_transverse("bad", swapped, 1, ql, qr, asdq, jac, eig)
