"""``import petclaw as pyclaw`` drop-in (src/petclaw/__init__.py:23-46): the same API as
``pyclaw`` with the State partitioned into y-slabs over the ranks of the current
torch.distributed job (one process per GPU), halo exchange over NCCL/NVLink and an
all-reduce(MAX) of the Courant number -- PetClaw's PETSc DMDA, re-hosted.

Launch with torchrun; if torch.distributed is not initialised the job is a single slab.
"""
import os
import sys as _sys

import torch
import torch.distributed as dist

import pyclaw_b200 as _impl
from pyclaw_b200 import *  # noqa: F401,F403
from pyclaw_b200 import (limiters, riemann, grid, solution, solver, clawpack, sharpclaw, controller, util)
from pyclaw_b200 import state as _state
from pyclaw_b200.parallel import SlabPartition, world


def init(backend=None):
    """Join the torchrun job (idempotent).  NCCL on GPUs, gloo on CPU-only hosts."""
    if dist.is_initialized() or 'RANK' not in os.environ:
        return world()
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    # a short collective timeout turns a mismatched exchange into a quick error instead of
    # a ten-minute hang
    import datetime
    dist.init_process_group(backend or ('nccl' if torch.cuda.is_available() else 'gloo'),
                            timeout=datetime.timedelta(seconds=int(os.environ.get('CLAWB200_DIST_TIMEOUT', 180))))
    return world()


class State(_state.State):
    """State on a slab partition (src/petclaw/state.py:136-167)."""

    def _make_partition(self, grid):
        part = getattr(grid, '_partition', None)
        if part is None:
            init()
            part = SlabPartition(grid)
            grid._partition = part
        return part


class Solution(_impl.solution.Solution):
    """Solution whose frames are read back onto the slab partition."""
    _state_class = State


class Controller(_impl.controller.Controller):
    """src/petclaw/controller.py:8-17: identical except for the default output format."""

    def __init__(self):
        super(Controller, self).__init__()
        self.output_format = 'petsc'


for _name in ('limiters', 'riemann', 'grid', 'solution', 'solver', 'clawpack', 'sharpclaw',
              'controller', 'util'):
    _sys.modules['petclaw.' + _name] = getattr(_impl, _name)
