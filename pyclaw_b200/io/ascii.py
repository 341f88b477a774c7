"""Classic Clawpack ascii frames (src/pyclaw/io/ascii.py:25-172 writer, :174-330 reader).

File layout, as the reference writes it:
  <prefix>.tNNNN   5 lines: t (%18.8e), meqn, nstates, maux, ndim (%5i), each followed by its name
  <prefix>.qNNNN   per state: grid_number, AMR_level, m<dim>..., <dim>low..., d<dim>..., blank
                   line, then one line per cell with meqn values in %18.8e, x fastest;
                   in 2-D a blank line after every row of constant y
  <prefix>.aNNNN   same with maux values per cell
"""
import logging
import os

import numpy as np

from ._common import global_array, is_writer, local_block

logger = logging.getLogger('io')


def _header(f, grid, n_global):
    f.write("%5i                  grid_number\n" % grid.gridno)
    f.write("%5i                  AMR_level\n" % grid.level)
    for dim, n in zip(grid.dimensions, n_global):
        f.write("%5i                  m%s\n" % (n, dim.name))
    for dim in grid.dimensions:
        f.write("%18.8e     %slow\n" % (dim.lower, dim.name))
    for dim in grid.dimensions:
        f.write("%18.8e     d%s\n" % (dim.d, dim.name))
    f.write("\n")


def _cells(f, a):
    """a is (ncomp, mx[, my]); one line per cell, x fastest, blank line per y row."""
    ncomp = a.shape[0]
    fmt = "%18.8e" * ncomp + "\n"
    if a.ndim == 2:
        f.write("".join(fmt % tuple(v) for v in a.T))
    elif a.ndim == 3:
        for j in range(a.shape[2]):
            f.write("".join(fmt % tuple(v) for v in a[:, :, j].T))
            f.write("\n")
    else:
        raise Exception("Dimension Exception in writing fort file.")


def write_ascii(solution, frame, path, file_prefix='fort', write_aux=False, options={}, write_p=False):
    tag = str(frame).zfill(4)
    blocks = []
    for state in solution.states:
        src = state.p if write_p else state.q
        q = global_array(state, src)
        aux = global_array(state, state.aux) if (state.maux > 0 and write_aux) else None
        blocks.append((state.grid, q, aux))
    if not is_writer():
        return
    try:
        with open(os.path.join(path, '%s.t%s' % (file_prefix, tag)), 'w') as f:
            f.write("%18.8e     time\n" % solution.t)
            f.write("%5i                  meqn\n" % blocks[0][1].shape[0])
            f.write("%5i                  nstates\n" % len(solution.states))
            f.write("%5i                  maux\n" % solution.maux)
            f.write("%5i                  ndim\n" % solution.ndim)
        with open(os.path.join(path, '%s.q%s' % (file_prefix, tag)), 'w') as qf:
            for grid, q, aux in blocks:
                _header(qf, grid, q.shape[1:])
                _cells(qf, q)
        if solution.maux > 0 and write_aux:
            with open(os.path.join(path, '%s.a%s' % (file_prefix, tag)), 'w') as af:
                for grid, q, aux in blocks:
                    _header(af, grid, aux.shape[1:])
                    _cells(af, aux)
    except IOError as e:
        logger.error("Error writing frame %s under %s: %s" % (frame, path, e))
        raise


def read_ascii_t(frame, path='./', file_prefix='fort'):
    """[t, meqn, nstates, maux, ndim] from <prefix>.tNNNN (ascii.py:332-377)."""
    fname = os.path.join(path, '%s.t' % file_prefix) + str(frame).zfill(4)
    with open(fname, 'r') as f:
        vals = [f.readline().split()[0] for _ in range(5)]
    return [float(vals[0])] + [int(v) for v in vals[1:]]


class _Lines(object):
    def __init__(self, f):
        self.f = f

    def value(self, kind=float):
        return kind(self.f.readline().split()[0])

    def numbers(self, count):
        out = []
        while len(out) < count:
            line = self.f.readline()
            if line == '':
                raise IOError("unexpected end of file: %d of %d values read" % (len(out), count))
            out.extend(line.split())
        if len(out) != count:
            raise IOError("cell lines do not add up to %d values" % count)
        return np.array(out, dtype=np.float64)


def _read_grid_header(L, ndim):
    from ..grid import Dimension, Grid
    gridno = L.value(int)
    level = L.value(int)
    n = [L.value(int) for _ in range(ndim)]
    lower = [L.value() for _ in range(ndim)]
    d = [L.value() for _ in range(ndim)]
    L.f.readline()
    names = ['x', 'y', 'z']
    grid = Grid([Dimension(names[i], lower[i], lower[i] + n[i] * d[i], n[i]) for i in range(ndim)])
    grid.gridno, grid.level = gridno, level
    return grid, n


def read_ascii(solution, frame, path='./', file_prefix='fort', read_aux=False, options={}):
    if frame < 0:
        raise IOError("Frame " + str(frame) + " does not exist ***")
    from .. import state as _state
    State = options.get('state_class', _state.State)
    t, meqn, nstates, maux, ndim = read_ascii_t(frame, path, file_prefix)
    if ndim > 2:
        raise NotImplementedError("3d still does not work!")
    tag = str(frame).zfill(4)
    with open(os.path.join(path, '%s.q' % file_prefix) + tag, 'r') as f:
        L = _Lines(f)
        for _ in range(nstates):
            grid, n = _read_grid_header(L, ndim)
            state = State(grid, meqn, maux)
            state.t = t
            if maux > 0:
                state.aux[...] = 0.
            vals = L.numbers(meqn * int(np.prod(n)))
            # file order: component fastest, then x, then y
            glob = vals.reshape(list(reversed(n)) + [meqn]).transpose(list(range(ndim, -1, -1)))
            state.q = np.ascontiguousarray(local_block(state, glob))
            solution.states.append(state)
    if maux > 0 and read_aux:
        fname = None
        for cand in (os.path.join(path, '%s.a' % file_prefix) + tag, os.path.join(path, '%s.a' % file_prefix)):
            if os.path.exists(cand):
                fname = cand
                break
        if fname is None:
            logger.info("Unable to open auxillary file for frame %s" % frame)
            return
        with open(fname, 'r') as f:
            L = _Lines(f)
            for state in solution.states:
                grid, n = _read_grid_header(L, ndim)
                if list(n) != [dim.n for dim in state.grid.dimensions]:
                    raise IOError("aux file grid does not match the q file grid")
                vals = L.numbers(maux * int(np.prod(n)))
                glob = vals.reshape(list(reversed(n)) + [maux]).transpose(list(range(ndim, -1, -1)))
                state.aux = np.ascontiguousarray(local_block(state, glob))
