"""PetClaw's restart format without PETSc (src/petclaw/io/petsc.py:32-232).

  <prefix>.pklNNNN   pickles: {t, meqn, nstates, maux, ndim, write_aux, aux_global}, then per
                     state {level, names, lower, n, d}
  <prefix>.ptcNNNN   one PETSc binary Vec per state: int32 class id 1211214, int32 length,
                     float64 values, all big-endian, in the DMDA's natural ordering
                     (component fastest, then x, then y) -- what ``gqVec.view(binary viewer)``
                     writes, so files are interchangeable with the reference's
  <prefix>_aux.ptc   same for aux (no frame number, as in the reference)
Only options['format'] == 'binary' is supported.
"""
import os
import pickle

import numpy as np

from ._common import global_array, is_writer, local_block

VEC_FILE_CLASSID = 1211214


def _write_vec(f, glob):
    flat = np.asarray(glob).reshape(-1, order='F')
    np.array([VEC_FILE_CLASSID, flat.size], dtype='>i4').tofile(f)
    flat.astype('>f8').tofile(f)


def _read_vec(f):
    head = np.fromfile(f, dtype='>i4', count=2)
    if head.size != 2 or head[0] != VEC_FILE_CLASSID:
        raise IOError("not a PETSc binary Vec")
    vals = np.fromfile(f, dtype='>f8', count=int(head[1]))
    if vals.size != head[1]:
        raise IOError("truncated PETSc binary Vec")
    return vals.astype(np.float64)


def write_petsc(solution, frame, path='./', file_prefix='claw', write_aux=False, options={}, write_p=False):
    opts = {'format': 'binary', 'clobber': True}
    opts.update(options)
    if opts['format'] != 'binary':
        raise IOError('format type %s not supported' % opts['format'])
    tag = str(frame).zfill(4)
    pickle_filename = os.path.join(path, '%s.pkl' % file_prefix) + tag
    viewer_filename = os.path.join(path, '%s.ptc' % file_prefix) + tag
    aux_filename = os.path.join(path, '%s_aux.ptc' % file_prefix)
    write_aux = bool(solution.maux > 0 and write_aux)
    if not opts['clobber']:
        for name in [pickle_filename, viewer_filename] + ([aux_filename] if write_aux else []):
            if os.path.exists(name):
                raise IOError('Cowardly refusing to clobber %s!' % name)
    blocks = []
    for state in solution.states:
        q = global_array(state, state.p if write_p else state.q)
        aux = global_array(state, state.aux) if write_aux else None
        blocks.append((state.grid, q, aux))
    if not is_writer():
        return
    with open(pickle_filename, 'wb') as pf, open(viewer_filename, 'wb') as vf:
        pickle.dump({'t': solution.t, 'meqn': solution.mp if write_p else solution.meqn,
                     'nstates': len(solution.states), 'maux': solution.maux, 'ndim': solution.ndim,
                     'write_aux': write_aux, 'aux_global': solution.aux_global}, pf)
        af = open(aux_filename, 'wb') if write_aux else None
        try:
            for grid, q, aux in blocks:
                pickle.dump({'level': grid.level, 'names': grid.name, 'lower': grid.lower,
                             'n': list(q.shape[1:]), 'd': grid.d}, pf)
                _write_vec(vf, q)
                if write_aux:
                    _write_vec(af, aux)
        finally:
            if af is not None:
                af.close()


def read_petsc(solution, frame, path='./', file_prefix='claw', read_aux=False, options={}):
    opts = {'format': 'binary'}
    opts.update(options)
    if opts['format'] != 'binary':
        raise IOError('format type %s not supported' % opts['format'])
    from ..grid import Dimension, Grid
    from .. import state as _state
    State = opts.get('state_class', _state.State)
    tag = str(frame).zfill(4)
    pickle_filename = os.path.join(path, '%s.pkl' % file_prefix) + tag
    viewer_filename = os.path.join(path, '%s.ptc' % file_prefix) + tag
    aux_filename = os.path.join(path, '%s_aux.ptc' % file_prefix)
    if frame < 0:
        raise IOError("Frame " + str(frame) + " does not exist ***")
    with open(pickle_filename, 'rb') as pf, open(viewer_filename, 'rb') as vf:
        head = pickle.load(pf)
        meqn, maux = head['meqn'], head['maux']
        read_aux = bool(read_aux and head.get('write_aux') and maux > 0 and os.path.exists(aux_filename))
        af = open(aux_filename, 'rb') if read_aux else None
        try:
            for _ in range(head['nstates']):
                g = pickle.load(pf)
                n = [int(v) for v in g['n']]
                dims = [Dimension(g['names'][i], g['lower'][i], g['lower'][i] + n[i] * g['d'][i], n[i])
                        for i in range(head['ndim'])]
                grid = Grid(dims)
                grid.level = g['level']
                state = State(grid, meqn, maux)
                state.t = head['t']
                state.aux_global = head['aux_global']
                glob = _read_vec(vf).reshape([meqn] + n, order='F')
                state.q = np.ascontiguousarray(local_block(state, glob))
                if maux > 0:
                    if read_aux:
                        gaux = _read_vec(af).reshape([maux] + n, order='F')
                        state.aux = np.ascontiguousarray(local_block(state, gaux))
                    else:
                        state.aux[...] = 0.
                solution.states.append(state)
        finally:
            if af is not None:
                af.close()
