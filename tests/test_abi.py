"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads without a GPU
and exports every symbol that include/clawb200.h declares; no compute is attempted."""
import ctypes
import os
import re

import pytest

from pyclaw_b200 import _lib, build as build_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "clawb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clawb200_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    build_mod.build()
    L = _lib.load()
    assert L.clawb200_version() >= 100


def test_every_declared_symbol_is_exported():
    L = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "symbol %s declared in clawb200.h is not exported" % n


def test_ctypes_signatures_cover_the_header():
    declared = set(_declared_symbols()) - {"clawb200_version", "clawb200_last_error", "clawb200_weno_table_doubles",
                                                "clawb200_step2_launches", "clawb200_step3_scratch_doubles"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_problem_struct_matches_header_layout():
    # offsets implied by the header's field order (natural alignment)
    P = _lib.Problem
    assert P.ndim.offset == 0 and P.mx.offset == 20 and P.dx.offset == 32
    assert P.method.offset == 48 and P.mthlim.offset == 76 and P.rp_id.offset == 108
    assert P.rp_params.offset == 112 and P.mstride.offset == 176 and P.pitch.offset == 184
    assert P.dt_dev.offset == 192 and P.weno_k.offset == 200 and P.weno_tab.offset == 208
    assert P.step2_mode.offset == 216 and ctypes.sizeof(P) == 224


def test_step2_launches_reports_the_single_pass_kernel():
    L = _lib.load()
    ac = _lib.make_problem(2, 3, 2, 2, 16, 16, 0.1, 0.1, _lib.RP_ACOUSTICS, [1.0, 4.0, 2.0, 2.0], [1, 2, 2, 0, 0, 0, 0], [4, 4])
    eu = _lib.make_problem(2, 5, 5, 2, 16, 16, 0.1, 0.1, _lib.RP_EULER5, [1.4, 0.4], [1, 2, 2, 0, 0, 0, 0], [4] * 5)
    assert L.clawb200_step2_launches(ctypes.byref(ac)) == 1
    assert L.clawb200_step2_launches(ctypes.byref(eu)) == 2
    ac.step2_mode = 1
    assert L.clawb200_step2_launches(ctypes.byref(ac)) == 2


def test_invalid_arguments_return_errors_without_a_gpu():
    # argument validation happens before any CUDA call, so it can be exercised on CPU
    P = _lib.make_problem(2, 5, 5, 2, 16, 16, 0.1, 0.1, _lib.RP_EULER5, [1.4, 0.4], None, [4] * 5)
    with pytest.raises(_lib.ClawB200Error, match="differ"):
        _lib.call("clawb200_step2", ctypes.byref(P), ctypes.c_void_p(8), ctypes.c_void_p(8), None, 0.1,
                  ctypes.c_void_p(8), None)
    P.method[5] = 1                      # mcapa without an aux array
    with pytest.raises(_lib.ClawB200Error, match="aux"):
        _lib.call("clawb200_step2", ctypes.byref(P), ctypes.c_void_p(8), ctypes.c_void_p(16), None, 0.1,
                  ctypes.c_void_p(8), None)
    P2 = _lib.make_problem(2, 4, 5, 2, 16, 16, 0.1, 0.1, _lib.RP_EULER5, [1.4, 0.4], None, [4] * 5)
    with pytest.raises(_lib.ClawB200Error, match="meqn"):
        _lib.call("clawb200_step2ds", ctypes.byref(P2), ctypes.c_void_p(8), ctypes.c_void_p(16), None, 0.1, 1,
                  ctypes.c_void_p(8), None)


def test_no_cpu_fallback():
    """On a host without CUDA the solvers must refuse to run rather than fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import pyclaw
    x = pyclaw.Dimension('x', 0., 1., 16)
    y = pyclaw.Dimension('y', 0., 1., 16)
    state = pyclaw.State(pyclaw.Grid([x, y]), 3)
    state.aux_global.update(rho=1., bulk=4., cc=2., zz=2.)
    solver = pyclaw.ClawSolver2D()
    solver.mwaves = 2
    for i in range(2):
        solver.bc_lower[i] = solver.bc_upper[i] = pyclaw.BC.outflow
    with pytest.raises(_lib.ClawB200Error, match="no CPU fallback"):
        solver.setup(pyclaw.Solution(state))
