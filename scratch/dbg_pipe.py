import sys, ctypes, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import problems
from oracle import pyclaw_oracle as po
from pyclaw_b200 import _lib
rp_id, params, meqn, mwaves, lim = 1, [1.0, 4.0, 2.0, 2.0], 3, 2, [4, 4]
mx, my, mbc = 24, 1500, 2
dx, dy, dt = 0.01, 0.013, 0.0011
q = problems.random_state("acoustics", (mx+4, my+4), 3)
method = [1, 2, 2, 0, 0, 0, 0]
P = _lib.make_problem(2, meqn, mwaves, mbc, mx, my, dx, dy, rp_id, params, method, lim)
qn_o = q.copy("F")
cfl_o = po.step2(rp_id, params, mbc, mx, my, q, qn_o, None, dx, dy, dt, method, lim)
qn_g = q.copy("F")
cfl_g = ctypes.c_double()
_lib.call("clawb200_step2_host", ctypes.byref(P), ctypes.c_void_p(q.ctypes.data), ctypes.c_void_p(qn_g.ctypes.data), None, dt, ctypes.byref(cfl_g))
d = np.abs(qn_g - qn_o)[:, 2:-2, 2:-2]
print('cfl', cfl_g.value, cfl_o, 'maxdiff', d.max())
bad = np.argwhere(d > 0)
print(len(bad), bad[:5], bad[-5:] if len(bad) else '')
print('rows with diffs', sorted(set(bad[:,2].tolist()))[:20])
