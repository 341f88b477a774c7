"""``import pyclaw`` drop-in: the reference's package name bound to pyclaw_b200."""
import sys as _sys

import pyclaw_b200 as _impl
from pyclaw_b200 import *  # noqa: F401,F403
from pyclaw_b200 import (limiters, riemann, grid, state, solution, solver, clawpack, sharpclaw,
                         controller, util)

for _name in ('limiters', 'riemann', 'grid', 'state', 'solution', 'solver', 'clawpack', 'sharpclaw',
              'controller', 'util'):
    _sys.modules['pyclaw.' + _name] = getattr(_impl, _name)
_sys.modules['pyclaw.limiters.tvd'] = _impl.limiters.tvd
