"""Partition independence on real GPUs: needs >= 2 visible devices, otherwise skipped."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("nranks", [2, 4])
def test_slab_partition_is_bit_identical(nranks):
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + nranks),
           os.path.join(HERE, "mp_partition_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
