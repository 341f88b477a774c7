"""Dimension and Grid (src/pyclaw/grid.py:36-143, 149-546), Python 3.

Coordinates stay on the host (numpy): they are set-up data, not part of the time step.
"""
import copy
import os

import numpy as np


def default_mapc2p(grid, x):
    return x


class Dimension(object):
    """grid.py:36-143.  ``Dimension(name, lower, upper, n)`` or ``Dimension(lower, upper, n)``."""

    def __init__(self, *args, **kargs):
        self.name = 'x'
        self.n = None
        self.lower = 0.0
        self.upper = 1.0
        self.units = None
        self._edge = None
        self._center = None
        if isinstance(args[0], float):
            self.lower, self.upper, self.n = float(args[0]), float(args[1]), int(args[2])
        elif isinstance(args[0], str):
            self.name = args[0]
            self.lower, self.upper, self.n = float(args[1]), float(args[2]), int(args[3])
        else:
            raise Exception("Invalid initializer for Dimension.")
        for (k, v) in kargs.items():
            setattr(self, k, v)
        # the part of the dimension owned by this process (grid.py:122-125; set by the
        # slab partition for multi-GPU runs)
        self.nstart = 0
        self.nend = self.n
        self.lowerg = self.lower

    @property
    def ng(self):
        """cells of this dimension owned by this process"""
        return self.nend - self.nstart

    @property
    def d(self):
        return (self.upper - self.lower) / float(self.n)

    @property
    def edge(self):
        if self._edge is None:
            # global index arithmetic: the coordinates of a cell do not depend on how the
            # dimension is partitioned (and equal the reference's lower + i*d in serial)
            self._edge = np.array([self.lower + (self.nstart + i) * self.d for i in range(self.ng + 1)])
        return self._edge

    @property
    def center(self):
        if self._center is None:
            self._center = np.array([self.lower + (self.nstart + i + 0.5) * self.d for i in range(self.ng)])
        return self._center

    def _set_range(self, nstart, nend):
        self.nstart, self.nend = nstart, nend
        self.lowerg = self.lower + nstart * self.d
        self._edge = self._center = None

    def __str__(self):
        output = "Dimension %s" % self.name
        if self.units:
            output += " (%s)" % self.units
        output += ":  (n,d,[lower,upper]) = (%s,%s,[%s,%s])" % (self.n, self.d, self.lower, self.upper)
        return output


class Grid(object):
    """grid.py:149-546."""

    def __init__(self, dimensions):
        self.level = 1
        self.gridno = 1
        self.mapc2p = default_mapc2p
        self.gauges = []
        self.gauge_files = []
        self.gauge_path = './_output/_gauges/'
        self._p_center = self._p_edge = self._c_center = self._c_edge = None
        if isinstance(dimensions, Dimension):
            dimensions = [dimensions]
        self._dimensions = []
        for dim in dimensions:
            self.add_dimension(dim)

    # ---- dimension bookkeeping ----
    def add_dimension(self, dimension):
        self._dimensions.append(dimension.name)
        setattr(self, dimension.name, dimension)

    def get_dim_attribute(self, attr):
        return [getattr(getattr(self, name), attr) for name in self._dimensions]

    ndim = property(lambda self: len(self._dimensions))
    dimensions = property(lambda self: [getattr(self, name) for name in self._dimensions])
    n = property(lambda self: self.get_dim_attribute('n'))
    ng = property(lambda self: self.get_dim_attribute('ng'))
    nstart = property(lambda self: self.get_dim_attribute('nstart'))
    nend = property(lambda self: self.get_dim_attribute('nend'))
    name = property(lambda self: self._dimensions)
    lower = property(lambda self: self.get_dim_attribute('lower'))
    lowerg = property(lambda self: self.get_dim_attribute('lowerg'))
    upper = property(lambda self: self.get_dim_attribute('upper'))
    d = property(lambda self: self.get_dim_attribute('d'))
    units = property(lambda self: self.get_dim_attribute('units'))
    center = property(lambda self: self.get_dim_attribute('center'))
    edge = property(lambda self: self.get_dim_attribute('edge'))

    @property
    def p_center(self):
        self.compute_p_center()
        return self._p_center

    @property
    def p_edge(self):
        self.compute_p_edge()
        return self._p_edge

    @property
    def c_center(self):
        self.compute_c_center()
        return self._c_center

    @property
    def c_edge(self):
        self.compute_c_edge()
        return self._c_edge

    def __str__(self):
        output = "Grid %s:\n" % self.gridno
        output += '\n  '.join((str(getattr(self, dim)) for dim in self._dimensions))
        return output + '\n'

    def is_valid(self):
        return True

    def __deepcopy__(self, memo={}):
        result = self.__class__(copy.deepcopy(self.dimensions))
        for attr in ('level', 'gridno', '_p_center', '_p_edge', '_c_center', '_c_edge'):
            setattr(result, attr, copy.deepcopy(getattr(self, attr)))
        result.mapc2p = self.mapc2p
        return result

    # ---- coordinate arrays (grid.py:364-512) ----
    def _mesh(self, which):
        arrays = self.get_dim_attribute(which)
        if self.ndim == 1:
            return [arrays[0]]
        return list(np.meshgrid(*arrays, indexing='ij'))

    def compute_c_center(self, recompute=False):
        if recompute or self._c_center is None:
            self._c_center = self._mesh('center')

    def compute_c_edge(self, recompute=False):
        if recompute or self._c_edge is None:
            self._c_edge = self._mesh('edge')

    def compute_p_center(self, recompute=False):
        if recompute or self._p_center is None:
            m = self._mesh('center')
            self._p_center = [self.mapc2p(self, m[0])] if self.ndim == 1 else self.mapc2p(self, m)

    def compute_p_edge(self, recompute=False):
        if recompute or self._p_edge is None:
            m = self._mesh('edge')
            self._p_edge = [self.mapc2p(self, m[0])] if self.ndim == 1 else self.mapc2p(self, m)

    # ---- gauges (grid.py:519-545) ----
    def add_gauges(self, gauge_coords):
        from numpy import floor
        if not os.path.exists(self.gauge_path):
            os.makedirs(self.gauge_path, exist_ok=True)
        for gauge in gauge_coords:
            gauge_ind = [int(floor(gauge[n] / self.d[n])) for n in range(self.ndim)]
            if all(self.nstart[n] <= gauge_ind[n] < self.nend[n] for n in range(self.ndim)):
                gauge_ind = [gauge_ind[n] - self.nstart[n] for n in range(self.ndim)]
                gauge_path = self.gauge_path + 'gauge' + '_'.join(str(coord) for coord in gauge) + '.txt'
                if os.path.isfile(gauge_path):
                    os.remove(gauge_path)
                self.gauges.append(list(gauge_ind))
                self.gauge_files.append(open(gauge_path, 'a'))
