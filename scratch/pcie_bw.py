"""Pinned host <-> device copy bandwidth, one direction and both at once (the e2e leg's ceiling)."""
import torch, time
n = 2_690_000_000 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_a = torch.empty(n, dtype=torch.float64, device="cuda")
d_b = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunks=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = n // chunks
    for k in range(chunks):
        sl = slice(k * c, (k + 1) * c)
        if h2d:
            with torch.cuda.stream(s1): d_a[sl].copy_(h_in[sl], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out[sl].copy_(d_b[sl], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t0
for name, a in (("H2D", (True, False)), ("D2H", (False, True)), ("both", (True, True)), ("both, 64 chunks", (True, True, 64))):
    run(*a); t = min(run(*a) for _ in range(3))
    print("%-16s %.1f ms  %.1f GB/s per direction" % (name, t * 1e3, n * 8 / t / 1e9))
