"""Solution (src/pyclaw/solution.py:35-456): a list of States with attribute forwarding,
``write`` / ``read`` through the formats of ``pyclaw_b200.io``."""
import os

from .grid import Grid, Dimension
from .state import State


def _fwd_state(name, settable=False):
    def fget(self):
        return getattr(self.states[0], name)

    def fset(self, value):
        for s in self.states:
            setattr(s, name, value)
    return property(fget, fset if settable else None)


def _fwd_grid(name):
    return property(lambda self: getattr(self.states[0].grid, name))


class Solution(object):
    def __init__(self, *arg, **kargs):
        self.states = []
        if len(arg) > 0:
            if isinstance(arg[0], State):
                self.states.append(arg[0])
            elif isinstance(arg[0], list) and len(arg[0]) > 0 and isinstance(arg[0][0], State):
                self.states = arg[0]
            elif isinstance(arg[0], (Grid, Dimension)) or isinstance(arg[0], list):
                raise Exception("Solution(grid) needs meqn in this snapshot's API: build a State first")
            elif isinstance(arg[0], int):
                # Solution(frame, path=..., format=..., file_prefix=..., read_aux=..., options=...)
                # (solution.py:168-199)
                frame = arg[0]
                defaults = {'path': './', 'format': 'ascii', 'file_prefix': None,
                            'read_aux': False, 'options': {}}
                for k in kargs:
                    if k not in defaults:
                        raise Exception("Invalid keyword argument %r" % k)
                defaults.update(kargs)
                self.read(frame, **defaults)
            else:
                raise Exception("Invalid argument list")

    state = property(lambda self: self.states[0])
    grid = property(lambda self: self.states[0].grid)
    t = _fwd_state('t', True)
    q = _fwd_state('q')
    p = _fwd_state('p')
    F = _fwd_state('F')
    aux = _fwd_state('aux', True)
    aux_global = _fwd_state('aux_global', True)
    meqn = _fwd_state('meqn')
    maux = _fwd_state('maux')
    mp = _fwd_state('mp', True)
    mF = _fwd_state('mF', True)
    mcapa = _fwd_state('mcapa', True)
    ndim = _fwd_grid('ndim')
    dimensions = _fwd_grid('dimensions')
    n = _fwd_grid('n')
    name = _fwd_grid('name')
    lower = _fwd_grid('lower')
    upper = _fwd_grid('upper')
    d = _fwd_grid('d')
    units = _fwd_grid('units')
    center = _fwd_grid('center')
    edge = _fwd_grid('edge')
    p_center = _fwd_grid('p_center')
    p_edge = _fwd_grid('p_edge')
    c_center = _fwd_grid('c_center')
    c_edge = _fwd_grid('c_edge')

    def is_valid(self):
        return all([state.is_valid() for state in self.states])

    def __str__(self):
        return "states:\n" + "".join(str(s) for s in self.states)

    def set_all_states(self, attr, value, overwrite=True):
        for state in self.states:
            if getattr(state, attr) is None or overwrite:
                setattr(state, attr, value)

    def write(self, frame, path='./', format='ascii', file_prefix=None, write_aux=False,
              options={}, write_p=False):
        """solution.py:356-404.  ``format`` is a name or a list of names of ``io.write_<name>``."""
        from . import io
        path = os.path.expandvars(os.path.expanduser(path))
        os.makedirs(path, exist_ok=True)
        for form in ([format] if isinstance(format, str) else list(format)):
            write_func = getattr(io, 'write_%s' % form, None)
            if write_func is None:
                raise IOError("unknown output format %r" % form)
            kw = {} if file_prefix is None else {'file_prefix': file_prefix}
            write_func(self, frame, path, write_aux=write_aux, options=options, write_p=write_p, **kw)

    def read(self, frame, path='./', format='ascii', file_prefix=None, read_aux=False, options={}):
        """solution.py:406-447"""
        from . import io
        path = os.path.expandvars(os.path.expanduser(path))
        read_func = getattr(io, 'read_%s' % format, None)
        if read_func is None:
            raise IOError("unknown input format %r" % format)
        opts = dict(options)
        opts.setdefault('state_class', self._state_class)
        kw = {} if file_prefix is None else {'file_prefix': file_prefix}
        read_func(self, frame, path, read_aux=read_aux, options=opts, **kw)

    _state_class = State
