"""Device array type handed to user code (state.q, state.aux, qbc in BC callbacks).

A thin torch.Tensor subclass so that the idioms of existing PyClaw scripts keep
working on device-resident data: ``state.q[0,:,:] = <numpy expression>``,
``numpy.loadtxt(...) - q[0]``, ``q.reshape([-1])``.
"""
import numpy as np
import torch


class ClawArray(torch.Tensor):
    def __setitem__(self, key, value):
        if isinstance(value, np.ndarray):
            value = torch.as_tensor(np.ascontiguousarray(value), dtype=self.dtype).to(self.device)
        return super().__setitem__(key, value)

    def __array__(self, dtype=None, copy=None):
        a = self.detach().as_subclass(torch.Tensor).cpu().numpy()
        return a if dtype is None else a.astype(dtype)

    # mixed numpy / device arithmetic (``numpy.loadtxt(...) - q[0]`` in the reference's tests)
    def _co(self, other):
        if isinstance(other, np.ndarray):
            return torch.as_tensor(np.ascontiguousarray(other), dtype=self.dtype).to(self.device)
        return other

    def __add__(self, o): return super().__add__(self._co(o))
    def __radd__(self, o): return super().__radd__(self._co(o))
    def __sub__(self, o): return super().__sub__(self._co(o))
    def __rsub__(self, o): return super().__rsub__(self._co(o))
    def __mul__(self, o): return super().__mul__(self._co(o))
    def __rmul__(self, o): return super().__rmul__(self._co(o))
    def __truediv__(self, o): return super().__truediv__(self._co(o))

    def __rtruediv__(self, o):
        o = self._co(o)
        if isinstance(o, torch.Tensor):
            return torch.div(o, self)
        return super().__rtruediv__(o)

    def copy(self, order="F"):
        """numpy-style copy (solver.py:660 ``state.q.copy('F')``)."""
        return self.clone()

    def flatten_f(self):
        """Fortran-order flattening, as ``ndarray.flatten('f')``."""
        return self.permute(*reversed(range(self.dim()))).reshape(-1)


def as_claw(t):
    return t.as_subclass(ClawArray)


def default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
