"""The oracle's own outputs on small seeded inputs are pinned (tests/golden/oracle_vectors.npz,
written by tests/golden/make_oracle_vectors.py): an accidental change of its arithmetic shows up
here on CPU, whatever the GPU path does."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_its_pinned_vectors():
    spec = importlib.util.spec_from_file_location("make_oracle_vectors", os.path.join(HERE, "golden", "make_oracle_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh = mod.flatten(mod.cases())
    pinned = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))
    assert sorted(fresh) == sorted(pinned.files) and len(fresh) >= 60
    for k in pinned.files:
        assert np.array_equal(fresh[k], pinned[k]), k
        assert np.isfinite(fresh[k]).all(), k
