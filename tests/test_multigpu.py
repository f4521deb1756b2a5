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


def test_two_rank_merge_and_novel_match_oracle():
    from kevlar_b200 import _lib
    if _lib.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', _free_port(), os.path.join(REPO, 'tests', '_mgpu_worker.py')]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    assert 'multi-GPU merge OK on 2 ranks' in res.stdout
