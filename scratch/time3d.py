"""Time the 3-D classic steps (unsplit step3 and the three step3ds sweeps) at n^3."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyclaw_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mbc = 2
N = n + 2 * mbc
dev = "cuda"
torch.manual_seed(0)
q = torch.rand(4, N, N, N, dtype=torch.float64, device=dev) - 0.5
aux = torch.ones(2, N, N, N, dtype=torch.float64, device=dev)
aux[:, :, :, N // 2:] = 2.0
qn = torch.empty_like(q)
cfl = torch.zeros(16, dtype=torch.float64, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
d = 2.0 / n
dt = 0.2 * d
for trans, name in ((22, "step3 trans_cor(22)"), (11, "step3 trans_inc(11)"), (0, "step3 no_trans(0)"), (-1, "3 x step3ds")):
    P = _lib.make_problem(3, 4, 2, mbc, n, n, d, d, _lib.RP_ACOUSTICS3D_VC, [], [1, 2, trans, 0, 0, 0, 2], [4, 4], maux=2,
                          pitch=N, mstride=N * N * N)
    if trans >= 0:
        S = torch.empty(_lib.load().clawb200_step3_scratch_doubles(ctypes.byref(P)), dtype=torch.float64, device=dev)
        run = lambda: _lib.call("clawb200_step3", ctypes.byref(P), n, d, p(q), p(qn), p(aux), dt, p(S), p(cfl), st)
    else:
        q2 = torch.empty_like(q)
        def run():
            _lib.call("clawb200_step3ds", ctypes.byref(P), n, d, p(q), p(qn), p(aux), dt, 1, p(cfl), st)
            _lib.call("clawb200_step3ds", ctypes.byref(P), n, d, p(qn), p(q2), p(aux), dt, 2, p(cfl), st)
            _lib.call("clawb200_step3ds", ctypes.byref(P), n, d, p(q2), p(qn), p(aux), dt, 3, p(cfl), st)
    for _ in range(2):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%-22s n=%d  %.3f ms/step  %.3e cell-updates/s" % (name, n, ms, n ** 3 / (ms * 1e-3)))
