"""Controller (src/pyclaw/controller.py:31-303): the output-time loop around
``solver.evolve_to_time``.  Frames are written through ``Solution.write`` in
``output_format`` ('ascii', 'petsc', a list of those, or None) and/or kept in memory with
``keep_copy``.  Plotting is outside the hot-path scope."""
import copy
import os

import numpy as np

from .solver import Solver
from .util import FrameCounter


class Controller(object):
    def __init__(self):
        self.viewable_attributes = ['rundir', 'outdir', 'overwrite', 'solver', 'keep_copy',
                                    'output_format', 'nout', 'outstyle', 'verbosity']
        self.xdir = os.getcwd()
        self.rundir = os.getcwd()
        self.outdir = os.getcwd() + '/_output'
        self.overwrite = True
        self.start_frame = 0
        self.solution = None
        self.solver = None
        self.keep_copy = False
        self.frames = []
        self.write_aux_init = False
        self.write_aux_always = False
        self.output_format = 'ascii'
        self.output_file_prefix = None
        self.outdir_p = './_output/_p'
        self.file_prefix_p = 'claw_p'
        self.output_options = {}
        self.tfinal = 1.0
        self.outstyle = 1
        self.verbosity = 0
        self.nout = 10
        self.out_times = np.linspace(0.0, self.tfinal, self.nout)
        self.nstepout = 1
        self.plotdata = None
        self.compute_p = None
        self.compute_F = None
        self.F_file_name = 'F'
        self.F_path = './_output/' + self.F_file_name + '.txt'

    def __str__(self):
        output = "Controller attributes:\n"
        for attr in self.viewable_attributes:
            output += "  %s = %s \n" % (attr, getattr(self, attr))
        return output

    def check_validity(self):
        if self.solver is None:
            raise Exception("No solver set in controller.")
        if not isinstance(self.solver, Solver):
            raise Exception("Solver is not of correct type.")
        if not self.solver.is_valid():
            raise Exception("The solver failed to initialize properly.")
        if not self.solution.is_valid():
            raise Exception("Initial solution is not valid.")

    def _write(self, frame, write_aux):
        """controller.py:242-259 / :277-291"""
        if self.output_format is None:
            return
        if self.compute_p is not None:
            self.compute_p(self.solution.state)
            self.solution.write(frame, self.outdir_p, self.output_format, self.file_prefix_p,
                                write_aux=False, options=self.output_options, write_p=True)
        self.solution.write(frame, self.outdir, self.output_format, self.output_file_prefix,
                            write_aux, self.output_options)

    def run(self):
        """controller.py:195-303"""
        frame = FrameCounter()
        frame.set_counter(self.start_frame)
        if self.keep_copy:
            self.frames = []
        self.solver.setup(self.solution)
        self.solver.dt = self.solver.dt_initial
        self.check_validity()
        self.solver.write_gauge_values(self.solution)
        if self.outstyle == 1:
            output_times = np.linspace(self.solution.t, self.tfinal, self.nout + 1)
        elif self.outstyle == 2:
            output_times = self.out_times
        elif self.outstyle == 3:
            output_times = np.ones((self.nout + 1))
        else:
            raise Exception("Invalid output style %s" % self.outstyle)
        if self.keep_copy:
            self.frames.append(copy.deepcopy(self.solution))
        if self.output_format is not None and os.path.exists(self.outdir) and self.overwrite == False:
            raise Exception("Refusing to overwrite existing output data. \
                 \nEither delete/move the directory or set controller.overwrite=True.")
        self._write(frame, self.write_aux_init)
        self.write_F('w')
        status = self.solver.status
        for t in output_times[1:]:
            if self.outstyle < 3:
                status = self.solver.evolve_to_time(self.solution, t)
            else:
                for n in range(self.nstepout):
                    status = self.solver.evolve_to_time(self.solution)
            frame.increment()
            if self.keep_copy:
                self.frames.append(copy.deepcopy(self.solution))
            self._write(frame, self.write_aux_always)
            self.write_F()
            for f in self.solution.state.grid.gauge_files:
                f.flush()
        self.solver.teardown()
        for f in self.solution.state.grid.gauge_files:
            f.close()
        return status

    def write_F(self, mode='a'):
        if self.compute_F is not None:
            self.compute_F(self.solution.state)
            F = [self.solution.state.sum_F(i) for i in range(self.solution.state.mF)]
            if self.is_proc_0():
                os.makedirs(os.path.dirname(self.F_path) or '.', exist_ok=True)
                with open(self.F_path, mode) as F_file:
                    F_file.write(str(self.solution.t) + ' ' + ' '.join(str(j) for j in F) + '\n')

    def is_proc_0(self):
        from .parallel import world
        return world()[0] == 0
