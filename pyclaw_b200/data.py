"""``pyclaw.Data``: attribute container with the Clawpack ``values =: name`` file format
(src/pyclaw/data.py:68-330).  Only the generic container is provided; the ``setrun.py``
generators (ClawData / AmrclawData / make_*datafile) belong to the classic Fortran drivers and
are outside the hot-path scope."""
import os


def _parse_value(value):
    """data.py:35-66: int, float (Fortran ``d`` exponents accepted), T/F booleans, lists."""
    value = value.strip()
    if not value:
        return None
    if len(value.split()) > 1:
        return [_parse_value(v) for v in value.split()]
    try:
        return int(value)
    except ValueError:
        pass
    try:
        return float(value.lower().replace('d', 'e')) if any(c.isdigit() for c in value) else float(value)
    except ValueError:
        pass
    if value.lower() in ('t', 'true', '.true.'):
        return True
    if value.lower() in ('f', 'false', '.false.'):
        return False
    return value


class Data(object):
    def __init__(self, data_files=[], attributes=None):
        self.__dict__['_attributes'] = []
        self.__dict__['_owners'] = {}
        if attributes:
            for a in attributes:
                self.add_attribute(a, None)
        if isinstance(data_files, str):
            data_files = [data_files]
        if data_files:
            self.read(data_files)

    def __setattr__(self, name, value):
        if not name.startswith('_') and name not in self._attributes:
            self._attributes.append(name)
        object.__setattr__(self, name, value)

    def __str__(self):
        return "\n".join("%s = %r" % (k, getattr(self, k)) for k in self._attributes)

    attributes = property(lambda self: list(self._attributes))

    def add_attribute(self, name, value=None, owner=None):
        setattr(self, name, value)
        self._owners[name] = owner

    def remove_attributes(self, arg_list):
        if isinstance(arg_list, str):
            arg_list = [arg_list]
        for name in arg_list:
            if name in self._attributes:
                self._attributes.remove(name)
                self._owners.pop(name, None)
                delattr(self, name)

    def has_attribute(self, name):
        return name in self._attributes

    def set_owner(self, name, owner):
        if name not in self._attributes:
            raise KeyError("No attribute named %s" % name)
        self._owners[name] = owner

    def get_owner(self, name):
        return self._owners.get(name, None)

    def iteritems(self):
        return [(k, getattr(self, k)) for k in self._attributes]

    items = iteritems

    def read(self, data_paths):
        if isinstance(data_paths, str):
            data_paths = [data_paths]
        for filename in data_paths:
            filename = os.path.abspath(filename)
            if not os.path.exists(filename):
                raise Exception("No such data file: %s" % filename)
            with open(filename) as f:
                for line in f:
                    if '=:' not in line:
                        continue
                    value, tail = line.split('=:')
                    self.add_attribute(tail.split()[0], _parse_value(value), filename)

    def write(self, data_files=None, supplementary_file=None):
        """One file per owner (or the single file given); lines ``value =: name``."""
        if isinstance(data_files, str):
            data_files = [data_files]
        groups = {}
        for name in self._attributes:
            owner = self._owners.get(name) or supplementary_file or (data_files[0] if data_files else None)
            if owner is None:
                raise Exception("attribute %s has no data file to be written to" % name)
            if data_files is not None and owner not in data_files and supplementary_file is None and len(data_files) == 1:
                owner = data_files[0]
            groups.setdefault(owner, []).append(name)
        for fname, names in groups.items():
            with open(fname, 'w') as f:
                for name in names:
                    v = getattr(self, name)
                    if isinstance(v, (list, tuple)):
                        s = " ".join(str(x) for x in v)
                    elif isinstance(v, bool):
                        s = 'T' if v else 'F'
                    else:
                        s = str(v)
                    f.write("%s =: %s\n" % (s.ljust(24), name))
