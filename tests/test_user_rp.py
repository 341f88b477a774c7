"""The Riemann-solver plugin seam on the CPU side: a user header compiles into a variant of the
library that exports the whole C ABI (nvcc cross-compiles without a GPU)."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "examples", "user_rp", "rp_kpp.cuh")


def test_user_header_builds_a_complete_library_variant():
    from pyclaw_b200 import build
    import test_abi
    lib = build.build_user(HEADER, "kpp")
    assert os.path.basename(lib) == "libclawb200_user_kpp.so" and os.path.exists(lib)
    L = ctypes.CDLL(lib)
    for sym in test_abi._declared_symbols():
        assert hasattr(L, sym), sym


def test_base_library_refuses_the_user_solver_id():
    """rp_id = CLAWB200_RP_USER in a build without a user header is an error, not a fallback
    (argument validation happens before any CUDA call, so this runs without a GPU)."""
    from pyclaw_b200 import _lib
    _lib.set_variant("strict")
    P = _lib.make_problem(2, 1, 2, 2, 8, 8, 1.0, 1.0, _lib.RP_USER, [])
    fake_in, fake_out = ctypes.c_void_p(4096), ctypes.c_void_p(8192)   # never dereferenced
    try:
        _lib.call("clawb200_step2ds", ctypes.byref(P), fake_in, fake_out, None, 0.1, 1, ctypes.c_void_p(64), None)
    except _lib.ClawB200Error as e:
        assert "no user Riemann solver" in str(e)
    else:
        raise AssertionError("expected CLAWB200_ERR_UNSUPPORTED")
