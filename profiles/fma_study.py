"""What does FMA contraction cost in accuracy, and what does it buy in time?

north_star asks for agreement with the reference's Fortran to a relative L-infinity of 1e-12
after N steps; the parity build is bit-exact (strict IEEE, -fmad=false).  This script runs
every BASELINE configuration to its reference end time with both builds of the library
(solver.arithmetic = 'strict' | 'fma', pyclaw_b200/build.py) and records

    rel_linf = max |q_fma - q_strict| / max |q_strict|        per conserved component, worst one

plus the distance of each build from the reference's golden files where one exists.
Run on a GPU box:   python profiles/fma_study.py > gpurun_out/fma_study.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden")


def rel_linf(a, b):
    a, b = np.asarray(a), np.asarray(b)
    worst = 0.0
    for m in range(a.shape[0]):
        scale = np.abs(b[m]).max()
        if scale > 0:
            worst = max(worst, float(np.abs(a[m] - b[m]).max() / scale))
    return worst


def main():
    import torch  # noqa: F401
    import test_gpu_golden as tg
    import pyclaw
    from pyclaw_b200.apps import shallow_sphere as app
    out = {}

    def both(fn):
        return {v: fn(v) for v in ("strict", "fma")}

    # config 2 golden variant: shock-bubble, dim-split + source term, 160x40, 170 steps
    r = both(lambda v: tg._shockbubble(arithmetic=v))
    gold = np.loadtxt(os.path.join(GOLD, "sb_density"))
    out["shockbubble_dimsplit_160x40"] = {
        "steps": {v: r[v][1]["numsteps"] for v in r},
        "rel_linf_fma_vs_strict": rel_linf(r["fma"][0], r["strict"][0]),
        "max_abs_density_vs_golden": {v: float(np.abs(np.asarray(r[v][0][0]) - gold).max()) for v in r},
        "reference_tolerance": "max abs < 1e-12 (test_examples.py:385-397)"}
    # the application's setting: unsplit, order_trans = 2
    r = both(lambda v: tg._shockbubble(arithmetic=v, dim_split=False, order_trans=2))
    out["shockbubble_unsplit_160x40"] = {
        "steps": {v: r[v][1]["numsteps"] for v in r},
        "rel_linf_fma_vs_strict": rel_linf(r["fma"][0], r["strict"][0])}

    # config 1: 2-D acoustics, classic (golden, dim-split) and unsplit
    gold = np.loadtxt(os.path.join(GOLD, "acoustics2D_solution"))
    r = both(lambda v: np.asarray(tg._acoustics2d("classic", arithmetic=v)))
    out["acoustics2d_classic_100x100"] = {
        "rel_linf_fma_vs_strict": float(np.abs(r["fma"] - r["strict"]).max() / np.abs(r["strict"]).max()),
        "frobenius_vs_golden": {v: float(np.linalg.norm(r[v] - gold)) for v in r},
        "reference_tolerance": "norm < 1e-14 (test_examples.py:239-254)"}
    r = both(lambda v: np.asarray(tg._acoustics2d("classic", arithmetic=v, dim_split=0, order_trans=2)))
    out["acoustics2d_unsplit_100x100"] = {
        "rel_linf_fma_vs_strict": float(np.abs(r["fma"] - r["strict"]).max() / np.abs(r["strict"]).max())}
    gold = np.loadtxt(os.path.join(GOLD, "ac_sc_solution"))
    r = both(lambda v: np.asarray(tg._acoustics2d("sharpclaw", arithmetic=v, lim_type=3)))
    out["acoustics2d_sharpclaw_100x100"] = {
        "rel_linf_fma_vs_strict": float(np.abs(r["fma"] - r["strict"]).max() / np.abs(r["strict"]).max()),
        "frobenius_vs_golden": {v: float(np.linalg.norm(r[v] - gold)) for v in r},
        "reference_tolerance": "norm < 1e-4 (test_examples.py:333-376)"}

    # config 3: shallow water, SharpClaw SSP33 (dam break, 60x60 to t = 1) and classic unsplit
    o = dict(time_integrator="SSP33", cfl_max=0.6, cfl_desired=0.5)
    r = both(lambda v: tg._shallow("sharpclaw", arithmetic=v, **o))
    out["shallow_sharpclaw_ssp33_60x60"] = {"rel_linf_fma_vs_strict": rel_linf(r["fma"], r["strict"])}
    # conditioning of that run: the STRICT build on initial data moved by one unit in the last
    # place (every other cell's depth times 1 + 2^-52).  If round-off of the input alone moves the
    # result by as much as the fma build does, the 1e-12 bar is a statement about bit-exactness
    # for this configuration, not about the quality of the arithmetic.
    import problems as _pb
    orig = _pb.shallow2d

    def perturbed(mx, my):
        pb = orig(mx, my)
        q = np.array(pb["q"], copy=True)
        q[0, ::2, ::2] *= (1.0 + 2.0 ** -52)
        pb = dict(pb)
        pb["q"] = q
        return pb
    _pb.shallow2d = perturbed
    try:
        rp = tg._shallow("sharpclaw", arithmetic="strict", **o)
    finally:
        _pb.shallow2d = orig
    out["shallow_sharpclaw_ssp33_60x60"]["rel_linf_strict_1ulp_input_perturbation"] = rel_linf(rp, r["strict"])
    r = both(lambda v: tg._shallow("classic", arithmetic=v, dim_split=0, order_trans=2))
    out["shallow_classic_unsplit_60x60"] = {"rel_linf_fma_vs_strict": rel_linf(r["fma"], r["strict"])}

    # config 4: shallow water on the sphere, 40x20 to t = 10 (764 steps)
    gold = np.loadtxt(os.path.join(GOLD, "swsphere_height"))

    def sphere(v):
        state, solver = app.setup(pyclaw)
        solver.arithmetic = v
        claw = pyclaw.Controller()
        claw.keep_copy, claw.output_format, claw.nout, claw.tfinal = True, None, 10, 10
        claw.solution, claw.solver = pyclaw.Solution(state), solver
        st = claw.run()
        return np.asarray(claw.frames[-1].state.q), st["numsteps"]
    r = both(sphere)
    out["shallow_sphere_40x20"] = {
        "steps": {v: r[v][1] for v in r},
        "rel_linf_fma_vs_strict": rel_linf(r["fma"][0], r["strict"][0]),
        "frobenius_height_vs_golden": {v: float(np.linalg.norm(r[v][0][0] - gold)) for v in r},
        "reference_tolerance": "norm < 1e-4 (test_examples.py:456-472)"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
