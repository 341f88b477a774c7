"""TVD limiter ids understood by the fused sweeps (philim.f:4-58 numbering).

Only the ids are kept (src/pyclaw/limiters/tvd.py:74-79); the limiter functions
themselves are device code in csrc/classic.cuh.
"""
minmod = 1
superbee = 2
vanleer = 3
MC = 4
beam_warming = 5
