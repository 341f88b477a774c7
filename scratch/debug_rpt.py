import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rp_properties as rpp, test_rp_properties as cpu, test_gpu_rp_properties as gpu
n = 8192
ql, qr, (sl, unl, unr) = rpp.euler_states(n, 1)
for ixy in (1, 2):
    l, r = ql.copy(), qr.copy()
    rpp._set_normal(l, ixy, sl, unl, 1.0); rpp._set_normal(r, ixy, sl, unr, 1.0)
    asdq = np.random.RandomState(7 + ixy).uniform(-1, 1, l.shape)
    for imp in (1, 2):
        g = gpu._transverse_for("euler")(ixy, l, r, imp, asdq)
        c = cpu._transverse_for("euler")(ixy, l, r, imp, asdq)
        d = np.abs(g[0] - c[0]) + np.abs(g[1] - c[1])
        bad = np.where(d.max(axis=0) > 0)[0]
        print("ixy", ixy, "imp", imp, "differing interfaces:", len(bad), bad[:10])
        for k in bad[:3]:
            print(" k", k, "l", l[:, k], "r", r[:, k], "\n  gpu bm", g[0][:, k], "\n  cpu bm", c[0][:, k])
    gs = gpu._solve_for("euler")(ixy, l, r); cs = cpu._solve_for("euler")(ixy, l, r)
    print("solve equal:", [bool(np.array_equal(a, b)) for a, b in zip(gs, cs)])
