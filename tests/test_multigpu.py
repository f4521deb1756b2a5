"""Multi-GPU parity: needs >= 2 GPUs on one box (skipped otherwise)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return str(s.getsockname()[1])


@pytest.mark.parametrize('world', [2, 4, 8])
def test_merge_and_novel_match_oracle(world):
    """Every merge strategy x 4 sketch types, plan B, and the shard-local novel scan against the
    single-process oracle, on `world` ranks (3 and 7 peers exercise slice_bounds and the one-pass
    all-reduce kernel beyond the 2-rank case)."""
    from kevlar_b200 import _lib
    if _lib.device_count() < world:
        pytest.skip('needs {} GPUs'.format(world))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr',
           '127.0.0.1', '--master-port', _free_port(), os.path.join(REPO, 'tests', '_mgpu_worker.py')]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    assert 'multi-GPU merge OK on {} ranks'.format(world) in res.stdout
