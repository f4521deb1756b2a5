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


@pytest.mark.parametrize('world', [2, 8])
def test_spanning_sketches_shared_memory_apply(world):
    """The spanning-sketch section of the worker again with 1024-bucket regions and the sparse-region
    shortcut off, so that the shared-memory apply with several source ranks does the work."""
    from kevlar_b200 import _lib
    if _lib.device_count() < world:
        pytest.skip('needs {} GPUs'.format(world))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr',
           '127.0.0.1', '--master-port', _free_port(), os.path.join(REPO, 'tests', '_mgpu_worker.py')]
    env = dict(os.environ, KV_MGPU_ONLY='span', KV_TILE_RB='10', KV_TILE_DIRECT_BELOW='0')
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    assert 'spanning sketches OK on {} ranks'.format(world) in res.stdout


@pytest.mark.parametrize('world', [2, 8])
def test_banded_chain_one_band_per_rank(world, tmp_path):
    """Config 5: bands spread over the ranks (band b on rank (b-1) mod world), rank 0 unbands and filters;
    all files equal to the reference chain's."""
    from kevlar_b200 import _lib
    if _lib.device_count() < world:
        pytest.skip('needs {} GPUs'.format(world))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr',
           '127.0.0.1', '--master-port', _free_port(), os.path.join(REPO, 'tests', '_bands_worker.py'), str(tmp_path)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    assert 'banded chain OK on {} ranks'.format(world) in res.stdout
