"""Helpers shared by the frame formats: host staging of (possibly partitioned) fields."""
import numpy as np
import torch

from ..parallel import world


def to_host(arr):
    if isinstance(arr, torch.Tensor):
        return arr.detach().cpu().numpy()
    return np.asarray(arr)


def global_array(state, arr):
    """The global (un-partitioned) array of a field given as this rank's interior block.
    Every rank takes part; every rank gets the result (frames are small next to the run)."""
    local = np.ascontiguousarray(to_host(arr))
    part = getattr(state, '_partition', None)
    if part is None or part.size == 1:
        return local
    import torch.distributed as dist
    parts = [None] * part.size
    dist.all_gather_object(parts, local, group=part.group)
    return np.concatenate(parts, axis=part.dim_index + 1)


def is_writer():
    return world()[0] == 0


def local_block(state, glob):
    """This rank's slab of a global array (inverse of global_array)."""
    part = getattr(state, '_partition', None)
    if part is None or part.size == 1:
        return glob
    dim = state.grid.dimensions[part.dim_index]
    idx = [slice(None)] * glob.ndim
    idx[part.dim_index + 1] = slice(dim.nstart, dim.nend)
    return glob[tuple(idx)]
