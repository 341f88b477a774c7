"""Frame formats: ascii fort.q/fort.t and the pickle + PETSc-binary
restart files; host-side logic only, runs without a GPU."""
import os
import pickle
import struct

import numpy as np
import pytest
import torch

import pyclaw


def _solution(mx=7, my=5, meqn=3, maux=2, ndim=2):
    if ndim == 2:
        grid = pyclaw.Grid([pyclaw.Dimension('x', -1.0, 1.0, mx), pyclaw.Dimension('y', 0.0, 0.5, my)])
        shape = (mx, my)
    else:
        grid = pyclaw.Grid([pyclaw.Dimension('x', -1.0, 1.0, mx)])
        shape = (mx,)
    state = pyclaw.State(grid, meqn, maux, device='cpu')
    rng = np.random.default_rng(3)
    state.q = rng.standard_normal((meqn,) + shape) * 10.0 ** rng.integers(-5, 5, (meqn,) + shape)
    if maux:
        state.aux = rng.standard_normal((maux,) + shape)
    state.t = 0.375
    state.aux_global = {'gamma': 1.4}
    return pyclaw.Solution(state)


def test_ascii_layout_matches_the_reference_writer(tmp_path):
    """Byte layout of fort.t / fort.q as src/pyclaw/io/ascii.py:62-117 produces it."""
    sol = _solution(mx=3, my=2, meqn=2, maux=0)
    sol.write(4, str(tmp_path), 'ascii')
    t = open(tmp_path / 'fort.t0004').read().split('\n')
    assert t[0] == "%18.8e     time" % 0.375
    assert t[1] == "    2                  meqn"
    assert t[2] == "    1                  nstates"
    assert t[3] == "    0                  maux"
    assert t[4] == "    2                  ndim"
    lines = open(tmp_path / 'fort.q0004').read().split('\n')
    assert lines[0] == "    1                  grid_number"
    assert lines[1] == "    1                  AMR_level"
    assert lines[2] == "    3                  mx"
    assert lines[3] == "    2                  my"
    assert lines[4] == "%18.8e     xlow" % -1.0
    assert lines[5] == "%18.8e     ylow" % 0.0
    assert lines[6] == "%18.8e     dx" % (2.0 / 3)
    assert lines[7] == "%18.8e     dy" % 0.25
    assert lines[8] == ""
    q = np.asarray(sol.q)
    k = 9
    for j in range(2):
        for i in range(3):
            assert lines[k] == "%18.8e%18.8e" % (q[0, i, j], q[1, i, j])
            k += 1
        assert lines[k] == ""
        k += 1


@pytest.mark.parametrize("ndim", [1, 2])
def test_ascii_round_trip(tmp_path, ndim):
    sol = _solution(ndim=ndim)
    sol.write(12, str(tmp_path), 'ascii', write_aux=True)
    back = pyclaw.Solution(12, path=str(tmp_path), format='ascii', read_aux=True)
    assert back.t == sol.t and back.meqn == sol.meqn and back.maux == sol.maux
    assert [d.n for d in back.dimensions] == [d.n for d in sol.dimensions]
    np.testing.assert_allclose(back.lower, sol.lower)
    np.testing.assert_allclose(back.d, sol.d, rtol=1e-8)
    # %18.8e keeps 9 significant digits
    np.testing.assert_allclose(np.asarray(back.q), np.asarray(sol.q), rtol=5e-9)
    np.testing.assert_allclose(np.asarray(back.aux), np.asarray(sol.aux), rtol=5e-9)
    # without read_aux the aux array comes back zeroed (ascii.py:256-262)
    back = pyclaw.Solution(12, path=str(tmp_path))
    assert float(np.abs(np.asarray(back.aux)).max()) == 0.0
    with pytest.raises(IOError):
        pyclaw.Solution(-1, path=str(tmp_path))
    with pytest.raises(IOError):
        pyclaw.Solution(99, path=str(tmp_path))


def test_petsc_files_are_petsc_binary_vecs_and_round_trip_exactly(tmp_path):
    sol = _solution()
    sol.write(3, str(tmp_path), 'petsc', write_aux=True)
    raw = open(tmp_path / 'claw.ptc0003', 'rb').read()
    classid, n = struct.unpack('>ii', raw[:8])
    assert classid == 1211214 and n == 3 * 7 * 5 and len(raw) == 8 + 8 * n
    vals = np.frombuffer(raw[8:], dtype='>f8')
    q = np.asarray(sol.q)
    # DMDA natural ordering: component fastest, then x, then y
    assert vals[1 + 3 * (2 + 7 * 4)] == q[1, 2, 4]
    with open(tmp_path / 'claw.pkl0003', 'rb') as f:
        head = pickle.load(f)
        g = pickle.load(f)
    assert head['meqn'] == 3 and head['maux'] == 2 and head['ndim'] == 2 and head['nstates'] == 1
    assert head['aux_global'] == {'gamma': 1.4} and head['t'] == 0.375
    assert list(g['n']) == [7, 5] and list(g['names']) == ['x', 'y']
    back = pyclaw.Solution(3, path=str(tmp_path), format='petsc', read_aux=True)
    assert np.array_equal(np.asarray(back.q), q)
    assert np.array_equal(np.asarray(back.aux), np.asarray(sol.aux))
    assert back.aux_global == {'gamma': 1.4} and back.t == 0.375
    with pytest.raises(IOError):
        sol.write(3, str(tmp_path), 'petsc', options={'clobber': False})
    with pytest.raises(IOError):
        sol.write(3, str(tmp_path), 'hdf5')


def test_write_p_and_format_list(tmp_path):
    sol = _solution(maux=0)
    sol.state.p = torch.ones((1, 7, 5), dtype=torch.float64) * 2.5
    sol.write(0, str(tmp_path), ['ascii', 'petsc'], file_prefix='claw_p', write_p=True)
    assert os.path.exists(tmp_path / 'claw_p.q0000') and os.path.exists(tmp_path / 'claw_p.ptc0000')
    back = pyclaw.Solution(0, path=str(tmp_path), format='petsc', file_prefix='claw_p')
    assert back.meqn == 1 and float(np.asarray(back.q).min()) == 2.5
