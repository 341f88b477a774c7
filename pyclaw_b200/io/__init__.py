"""Frame output / restart formats.  Host-side only: q is staged from
HBM to the host once per output frame, never inside the time-step path.

    ascii : fort.tNNNN / fort.qNNNN / fort.aNNNN       (src/pyclaw/io/ascii.py)
    petsc : claw.pklNNNN + claw.ptcNNNN (PETSc binary Vec, natural ordering)
                                                       (src/petclaw/io/petsc.py)
"""
from .ascii import write_ascii, read_ascii, read_ascii_t
from .petsc import write_petsc, read_petsc

__all__ = ['write_ascii', 'read_ascii', 'read_ascii_t', 'write_petsc', 'read_petsc']
