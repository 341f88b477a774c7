"""Courant number holder (src/pyclaw/cfl.py:4-25)."""


class CFL(object):
    def __init__(self, global_max):
        self._global_max = global_max

    def get_global_max(self):
        return self._global_max

    def get_cached_max(self):
        return self._global_max

    def set_local_max(self, new_local_max):
        self._global_max = new_local_max

    def update_global_max(self, new_local_max):
        # replaces, does not accumulate (cfl.py:24-25)
        self._global_max = new_local_max
