#!/usr/bin/env python
"""Writes tests/golden/oracle_vectors.npz: outputs of the CPU oracle (oracle/claw_oracle.c) on
small seeded inputs, one entry per routine / Riemann solver.  tests/test_oracle_vectors.py
compares a fresh oracle build with them, so that an accidental change of the oracle's
arithmetic is caught on CPU, independently of the GPU parity tests.

    python tests/golden/make_oracle_vectors.py          # regenerate (only after a deliberate change)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems  # noqa: E402
from oracle import pyclaw_oracle as po  # noqa: E402


def cases():
    """name -> (callable returning a dict of arrays)"""
    out = {}
    mbc, mx, my = 2, 19, 13
    dx, dy, dt = 0.01, 0.013, 0.0011
    rps2 = {"acoustics": (1, [1.0, 4.0, 2.0, 2.0], [4, 4]), "advection": (2, [0.7, -0.4], [3]),
            "euler": (3, [1.4, 0.4], [4, 4, 4, 4, 2]), "shallow": (4, [1.0], [4, 1, 2])}
    for name, (rp, params, lim) in rps2.items():
        q = problems.random_state(name, (mx + 2 * mbc, my + 2 * mbc), seed=11)
        for trans in (-1, 0, 2):
            method = [1, 2, trans, 0, 0, 0, 0]
            if trans < 0:
                qn = q.copy("F")
                c1 = po.step2ds(rp, params, mbc, mx, my, q, qn, None, dx, dy, dt, method, lim, 1)
                qn2 = qn.copy("F")
                c2 = po.step2ds(rp, params, mbc, mx, my, qn, qn2, None, dx, dy, dt, method, lim, 2)
                out["step2ds_%s" % name] = dict(q=qn2, cfl=np.array([c1, c2]))
            else:
                qn = q.copy("F")
                c = po.step2(rp, params, mbc, mx, my, q, qn, None, dx, dy, dt, method, lim)
                out["step2_%s_trans%d" % (name, trans)] = dict(q=qn[:, mbc:-mbc, mbc:-mbc], cfl=np.array([c]))
        qs = problems.smooth_state(name, (mx + 6, my + 6), seed=5)
        for variant in (0, 1, 2):
            dq, c = po.sc_flux2(rp, params, len(lim), 3, mx, my, qs, dx, dy, dt, variant)
            out["sc_flux2_%s_v%d" % (name, variant)] = dict(dq=dq[:, 3:-3, 3:-3], cfl=np.array([c]))
    # capacity function
    aux = np.asfortranarray(np.random.RandomState(7).uniform(0.5, 1.5, (1, mx + 2 * mbc, my + 2 * mbc)))
    q = problems.random_state("acoustics", (mx + 2 * mbc, my + 2 * mbc), seed=3)
    qn = q.copy("F")
    c = po.step2(1, [1.0, 4.0, 2.0, 2.0], mbc, mx, my, q, qn, aux, dx, dy, dt, [1, 2, 2, 0, 0, 1, 1], [4, 4])
    out["step2_acoustics_capa"] = dict(q=qn[:, mbc:-mbc, mbc:-mbc], cfl=np.array([c]))
    # 1-D
    for name, rp, params, meqn, lim in (("acoustics", 1, [1.0, 1.0, 1.0, 1.0], 2, [4, 4]), ("advection", 2, [0.7], 1, [3]),
                                        ("shallow", 4, [1.0], 2, [4, 4]), ("burgers", po.RP_BURGERS, [], 1, [3]),
                                        ("euler", po.RP_EULER1D, [1.4, 0.4], 3, [4, 4, 4])):
        n = 40
        if name == "burgers":
            q = np.asfortranarray(np.random.RandomState(2).uniform(-1, 1, (1, n + 4)))
        elif name == "euler":
            q5 = problems.random_state("euler", (n + 4,), 4)
            q = np.asfortranarray(np.stack([q5[0], q5[1], q5[3] - 0.5 * q5[2] ** 2 / q5[0]]))
        elif name == "shallow":
            q = np.asfortranarray(problems.random_state("shallow", (n + 4,), 4)[:2])
        else:
            q = problems.random_state(name, (n + 4,), seed=4)
        qo = q.copy("F")
        c = po.step1(rp, params, 2, n, qo, None, 1.0 / n, 0.2 / n, [1, 2, 0, 0, 0, 0, 0], lim)
        out["step1_%s" % name] = dict(q=qo[:, 2:-2], cfl=np.array([c]))
    # f-waves (linear stress law: no libm), p-system with transverse solver
    rng = np.random.RandomState(9)
    n = 30
    q = np.asfortranarray(rng.uniform(-0.3, 0.3, (2, n + 4)))
    aux = np.asfortranarray(np.stack([rng.choice([1.0, 4.0], n + 4), rng.choice([1.0, 4.0], n + 4), np.zeros(n + 4)]))
    qo = q.copy("F")
    c = po.step1(po.RP_NEL_FWAVE, [1.0], 2, n, qo, aux, 0.1, 0.02, [1, 2, 0, 0, 0, 0, 3], [4, 4])
    out["step1_fwave_elasticity"] = dict(q=qo[:, 2:-2], cfl=np.array([c]))
    pad = (mx + 4, my + 4)
    q = np.asfortranarray(rng.uniform(-0.3, 0.3, (3,) + pad))
    aux = np.asfortranarray(np.stack([rng.choice([1.0, 4.0], pad), rng.choice([1.0, 4.0], pad), np.ones(pad), q[0]]))
    qn = q.copy("F")
    c = po.step2(po.RP_PSYSTEM, [], 2, mx, my, q, qn, aux, 0.05, 0.04, 0.008, [1, 2, 2, 0, 0, 0, 4], [2, 2])
    out["step2_fwave_psystem"] = dict(q=qn[:, 2:-2, 2:-2], cfl=np.array([c]))
    # sphere
    pb = problems.sphere_problem(16, 8)
    qbc = np.zeros((4, 20, 12), order="F")
    qbc[:, 2:-2, 2:-2] = pb["q"]
    po.fill_bcs(qbc, 2, [po.BC_PERIODIC, po.BC_CUSTOM], [po.BC_PERIODIC, po.BC_CUSTOM],
                problems.sphere_qbc_lower_y, problems.sphere_qbc_upper_y)
    qn = qbc.copy("F")
    dxs, dys = pb["d"]
    c = po.step2(po.RP_SPHERE, [problems.SPHERE_G, dxs, dys], 2, 16, 8, qbc, qn, pb["auxbc_full"], dxs, dys,
                 0.4 * dxs / 4.0, [1, 2, 2, 0, 0, 1, 16], [4, 4, 4])
    out["step2_sphere"] = dict(q=qn[:, 2:-2, 2:-2], cfl=np.array([c]))
    # 3-D dimensional splitting
    pad3 = (9 + 4, 7 + 4, 5 + 4)
    q = np.asfortranarray(rng.uniform(-1, 1, (4,) + pad3))
    aux = np.asfortranarray(np.stack([rng.choice([1.0, 2.0, 3.5], pad3), rng.choice([1.0, 2.0], pad3)]))
    for idir in (1, 2, 3):
        qn = q.copy("F")
        c = po.step3ds(po.RP_ACOUSTICS3D_VC, [], 2, 9, 7, 5, q, qn, aux, 0.02, 0.025, 0.03, 0.004,
                       [1, 2, -1, 0, 0, 0, 2], [4, 4], idir)
        out["step3ds_dir%d" % idir] = dict(q=qn, cfl=np.array([c]))
    return out


def flatten(cs):
    return {"%s/%s" % (k, kk): v for k, d in cs.items() for kk, v in d.items()}


if __name__ == "__main__":
    po.build()
    flat = flatten(cases())
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **flat)
    print("wrote %d arrays, %d bytes" % (len(flat), os.path.getsize(os.path.join(HERE, "oracle_vectors.npz"))))
