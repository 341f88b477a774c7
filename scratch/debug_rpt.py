import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rp_properties as rpp, test_rp_properties as cpu, test_gpu_rp_properties as gpu
np.set_printoptions(precision=17, linewidth=200)
n = 4096
for name, (ql, qr) in (("euler", rpp.euler_states(n, 11)[:2]), ("shallow", rpp.shallow_states(n, 12))):
    for ixy in (1, 2):
        for rep in range(2):
            g = gpu._solve_for(name)(ixy, ql, qr); c = cpu._solve_for(name)(ixy, ql, qr)
            for gg, cc, what in zip(g, c, ("wave", "s", "amdq", "apdq")):
                d = np.abs(gg - cc)
                d = np.where(np.isnan(gg) & np.isnan(cc), 0.0, d)
                bad = np.unique(np.where(~(d == 0))[-1])
                print(name, "ixy", ixy, "rep", rep, what, "differing interfaces:", len(bad), bad[:8], "nan in gpu", int(np.isnan(gg).sum()), "nan in cpu", int(np.isnan(cc).sum()))
                if len(bad) and what == "wave" and rep == 0:
                    k = bad[0]
                    print("  k", k, "\n  l", ql[:, k], "\n  r", qr[:, k], "\n  gpu", gg[..., k].ravel(), "\n  cpu", cc[..., k].ravel(), "\n  s gpu", g[1][:, k], "cpu", c[1][:, k])
