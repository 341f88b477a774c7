"""Riemann-solver descriptors.

The reference bakes the Riemann solver into classic2.so / sharpclaw2.so at link time
(RP_SOURCE in each application's Makefile, e.g. apps/euler/2d/shockbubble/Makefile:3), so
its scripts never name the solver in Python.  Here ``solver.rp`` may be set to one of
these descriptors (or its name); if it is left unset the solver is inferred from the
keys of ``state.aux_global`` (the cparam common block the script fills).
"""
from .._lib import (RP_ACOUSTICS, RP_ADVECTION, RP_EULER5, RP_SHALLOW, RP_SPHERE, RP_NEL_FWAVE,
                    RP_PSYSTEM, RP_ACOUSTICS3D_VC, RP_VC_ACOUSTICS, RP_BURGERS, RP_ADVECTION_COLOR,
                    RP_VC_ADVECTION, RP_EULER1D)


class RiemannSolver(object):
    def __init__(self, name, rp_id, meqn, mwaves, param_names, ndims, optional=(), defaults=None,
                 fwave=False, maux=0):
        self.name, self.rp_id, self.param_names, self.ndims = name, rp_id, param_names, ndims
        self._meqn = meqn
        self._mwaves = mwaves
        self.mwaves = self.nwaves(max(ndims))
        self.optional = set(optional)
        self.defaults = dict(defaults or {})
        self.fwave = fwave      # returns f-waves: pairs with the classic*fw modules (clawpack.py:222)
        self.maux = maux        # aux components the solver reads

    def meqn(self, ndim):
        return self._meqn(ndim) if callable(self._meqn) else self._meqn

    def nwaves(self, ndim):
        return self._mwaves(ndim) if callable(self._mwaves) else self._mwaves

    def params(self, aux_global):
        missing = [k for k in self.param_names if k not in aux_global and k not in self.optional]
        if missing:
            # state.py:154-158: every cparam variable must be present in aux_global
            raise Exception("Some required value(s) in the cparam common block in the Riemann "
                            "solver have not been set in aux_global: %s" % missing)
        return [float(aux_global.get(k, self.defaults.get(k, 0.0))) for k in self.param_names]

    def __repr__(self):
        return "<riemann %s>" % self.name


acoustics = RiemannSolver("acoustics", RP_ACOUSTICS, lambda ndim: ndim + 1, 2, ["rho", "bulk", "cc", "zz"], (1, 2))
advection = RiemannSolver("advection", RP_ADVECTION, 1, 1, ["u", "v"], (1, 2), optional=["v"])
euler_5wave = RiemannSolver("euler_5wave", RP_EULER5, 5, 5, ["gamma", "gamma1"], (2,))
shallow_roe_with_efix = RiemannSolver("shallow_roe_with_efix", RP_SHALLOW, lambda ndim: ndim + 1,
                                      lambda ndim: ndim + 1, ["grav"], (1, 2))

# shallow water on the sphere (apps/shallow-sphere): `g` is the reference's common /sw/ g;
# dxcom / dycom (common /comxyt/) default to the grid spacing
shallow_sphere = RiemannSolver("shallow_sphere", RP_SPHERE, 4, 3, ["g", "dxcom", "dycom"], (2,),
                               optional=["dxcom", "dycom"])

# f-wave solvers (solver.fwave = True).  Stress law 1: sigma = K eps; 2: sigma = exp(K eps) - 1.
#   1-D: aux = {rho, K}     (apps/elasticity/1d/stegoton/stegoton.py:17-27), law from
#        aux_global['stress_law'] (default 2, the stegoton's)
#   2-D: aux = {rho, E, stress law, copy of eps}   (test/psystem/psystem.py:35-86)
nonlinear_elasticity_fwave = RiemannSolver("nonlinear_elasticity_fwave", RP_NEL_FWAVE, 2, 2, ["stress_law"],
                                           (1,), optional=["stress_law"], defaults={"stress_law": 2.0},
                                           fwave=True, maux=2)
psystem = RiemannSolver("psystem", RP_PSYSTEM, 3, 2, [], (2,), fwave=True, maux=4)

# 3-D acoustics in a heterogeneous medium (test/acoustics/3d): aux = {impedance, sound speed}
vc_acoustics_3d = RiemannSolver("vc_acoustics_3d", RP_ACOUSTICS3D_VC, 4, 2, [], (3,), maux=2)

# further solvers of the reference's applications (set solver.rp to use them)
vc_acoustics = RiemannSolver("vc_acoustics", RP_VC_ACOUSTICS, 3, 2, [], (2,), maux=2)          # aux {rho, c}
burgers = RiemannSolver("burgers", RP_BURGERS, 1, 1, [], (1,))
advection_color = RiemannSolver("advection_color", RP_ADVECTION_COLOR, 1, 1, [], (1,), maux=1)  # aux {u}
vc_advection = RiemannSolver("vc_advection", RP_VC_ADVECTION, 1, 1, [], (2,), maux=2)           # aux {u, v[, capa]}
euler_with_efix = RiemannSolver("euler_with_efix", RP_EULER1D, 3, 3, ["gamma", "gamma1"], (1,))

_BY_NAME = {s.name: s for s in (vc_acoustics, burgers, advection_color, vc_advection, euler_with_efix,
                                vc_acoustics_3d, acoustics, advection, euler_5wave, shallow_roe_with_efix, shallow_sphere,
                                nonlinear_elasticity_fwave, psystem)}
_BY_NAME.update({"euler": euler_5wave, "shallow": shallow_roe_with_efix})


class _Module(object):
    """Mimics ``riemann.rp_acoustics.rp_acoustics_1d`` style access."""

    def __init__(self, solver, names):
        self.mwaves = solver.mwaves
        for n in names:
            setattr(self, n, solver)


rp_acoustics = _Module(acoustics, ["rp_acoustics_1d", "rp_acoustics_2d"])
rp_advection = _Module(advection, ["rp_advection_1d", "rp_advection_2d"])
rp_euler = _Module(euler_5wave, ["rp_euler_5wave_2d"])
rp_shallow = _Module(shallow_roe_with_efix, ["rp_shallow_roe_with_efix_2d"])


def resolve(rp, aux_global, ndim, fwave=False):
    if isinstance(rp, RiemannSolver):
        return rp
    if isinstance(rp, str):
        if rp not in _BY_NAME:
            raise Exception("Unknown Riemann solver %r" % rp)
        return _BY_NAME[rp]
    if rp is not None:
        raise NotImplementedError("Python Riemann solvers are not supported: there is no CPU "
                                  "path; set solver.rp to a pyclaw.riemann descriptor")
    found = _infer(aux_global, ndim, fwave)
    # The reference binds the solver at link time (RP_SOURCE in the application's Makefile); a
    # guess from the cparam names must be visible, so that a script whose aux_global happens to
    # match another physics does not run the wrong solver silently.
    import logging
    logging.getLogger('evolve').warning(
        "solver.rp is not set: using Riemann solver %r inferred from aux_global keys %s "
        "(set solver.rp = pyclaw.riemann.<name> to choose explicitly)", found.name, sorted(aux_global.keys()))
    return found


def _infer(aux_global, ndim, fwave):
    if ndim == 3:
        return vc_acoustics_3d       # the only 3-D solver the reference's applications link
    if fwave:
        # the only f-wave solvers the reference's applications link (stegoton, psystem)
        return nonlinear_elasticity_fwave if ndim == 1 else psystem
    keys = set(aux_global.keys())
    if {"rho", "bulk", "cc", "zz"} <= keys:
        return acoustics
    if {"gamma", "gamma1"} <= keys:
        return euler_with_efix if ndim == 1 else euler_5wave
    if "grav" in keys:
        return shallow_roe_with_efix
    if "u" in keys:
        return advection
    raise Exception("Cannot infer the Riemann solver from aux_global keys %s; set solver.rp" % sorted(keys))
