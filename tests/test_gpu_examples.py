"""examples/apps.py: every application of the reference runs through ``import pyclaw`` at a
reduced size, stays finite, and honours the invariants its equations have."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))


def _frames(claw):
    return np.asarray(claw.frames[0].q), np.asarray(claw.frames[-1].q)


@pytest.mark.parametrize("name,kwargs,conserved", [
    ("acoustics1d", dict(mx=100), [0, 1]),
    ("acoustics1d", dict(mx=100, solver_type='sharpclaw', weno_order=9), [0, 1]),
    ("acoustics2d", dict(mx=60, my=60, tfinal=0.06), [0]),
    ("acoustics2d", dict(mx=60, my=60, tfinal=0.06, dim_split=False), [0]),
    ("vc_acoustics2d", dict(mx=60, my=60, tfinal=0.2), []),
    ("shockbubble", dict(mx=80, my=20, tfinal=0.05), []),
    ("shallow1d", dict(mx=200, tfinal=1.0), [0]),
    ("shallow1d", dict(mx=200, tfinal=1.0, solver_type='sharpclaw'), [0]),
    ("shallow2d", dict(mx=50, my=50, tfinal=0.5), [0]),
    ("shallow2d", dict(mx=50, my=50, tfinal=0.5, solver_type='sharpclaw'), [0]),
    ("shallow_sphere", dict(mx=40, my=20, tfinal=0.5), []),
    ("stegoton", dict(layers=30, tfinal=10.0), [0, 1]),
    ("stegoton", dict(layers=30, tfinal=10.0, solver_type='sharpclaw'), [0, 1]),
    ("psystem", dict(cells_per_layer=8, tfinal=0.3), []),
    ("acoustics3d", dict(mx=64, my=4, mz=4, tfinal=0.5), [0]),
    ("burgers", dict(mx=200, tfinal=0.3), [0]),
    ("burgers", dict(mx=200, tfinal=0.3, solver_type='sharpclaw'), [0]),
    ("wcblast", dict(mx=300, tfinal=0.01), [0, 2]),
    ("vc_advection1d", dict(mx=100, tfinal=0.3), []),
    ("annulus", dict(mx=20, my=60, tfinal=0.3), []),
])
def test_application_runs(name, kwargs, conserved):
    import apps
    claw = apps.APPS[name](**kwargs)
    q0, q1 = _frames(claw)
    assert np.isfinite(q1).all()
    assert abs(claw.frames[-1].t - kwargs.get('tfinal', claw.tfinal)) < 1e-12
    assert np.abs(q1 - q0).max() > 1e-6                      # something happened
    for m in conserved:                                      # closed / periodic domains, before waves leave
        s0, s1 = q0[m].sum(), q1[m].sum()
        assert abs(s1 - s0) <= 1e-9 * max(1.0, np.abs(q0[m]).sum()), (name, m, s0, s1)


def test_annulus_rotation_preserves_the_capacity_weighted_mass():
    import apps
    claw = apps.annulus(mx=20, my=60, tfinal=0.3)
    kappa = np.asarray(claw.solution.state.aux)[2]
    q0, q1 = _frames(claw)
    # solid-body rotation: nothing crosses the inner / outer radius
    assert abs((kappa * q1[0]).sum() - (kappa * q0[0]).sum()) < 1e-10 * (kappa * q0[0]).sum() + 1e-12
