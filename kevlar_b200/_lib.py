"""ctypes binding of libkvsketch.so (include/kvsketch.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C kevlar_b200/csrc``.
There is no CPU fallback: a missing library or a missing GPU raises, loudly, at the first
call that needs it.
"""
import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_int, c_int64, c_uint64, c_void_p

import numpy as np

HASH_MURMUR = 0
HASH_TWOBIT = 1
MEM_HOST = 0
MEM_DEVICE = 1
MAX_TABLES = 8
MAX_SAMPLES = 16

KV_OK = 0
KV_EINVAL = -1
KV_EIO = -2
KV_ENOMEM = -3
KV_ECUDA = -4
KV_ENODEVICE = -5
KV_EOVERFLOW = -6
KV_ESTATE = -7

READ_SKIPPED = 1
READ_DISCARDED = 2

HIT_DTYPE = np.dtype([('read', '<u4'), ('offset', '<u4'), ('abund', 'u1', (MAX_SAMPLES,))])
assert HIT_DTYPE.itemsize == 24

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get('KV_LIB_PATH') or os.path.join(_HERE, 'libkvsketch.so')   # KV_LIB_PATH: experiment builds

# every symbol include/kvsketch.h declares: (name, restype, argtypes)
_P = c_void_p
SYMBOLS = [
    ('kv_last_error', c_char_p, []),
    ('kv_abi_version', c_int, []),
    ('kv_device_count', c_int, [POINTER(c_int)]),
    ('kv_primes_below', c_int, [c_uint64, c_int, POINTER(c_uint64)]),
    ('kv_sketch_create', c_int, [c_int, c_int, c_int, c_int, POINTER(c_uint64), c_int, POINTER(_P)]),
    ('kv_sketch_destroy', c_int, [_P]),
    ('kv_sketch_clear', c_int, [_P]),
    ('kv_sketch_load', c_int, [c_char_p, c_int, c_int, c_int, POINTER(_P)]),
    ('kv_sketch_save', c_int, [_P, c_char_p]),
    ('kv_sketch_info', c_int, [_P, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int),
                               POINTER(c_uint64), POINTER(c_int)]),
    ('kv_sketch_stats', c_int, [_P, POINTER(c_uint64), POINTER(c_uint64), POINTER(c_int)]),
    ('kv_sketch_set_unique_tracking', c_int, [_P, c_int]),
    ('kv_sketch_table', c_int, [_P, c_int, POINTER(_P), POINTER(c_uint64)]),
    ('kv_sketch_flat', c_int, [_P, POINTER(_P), POINTER(c_uint64)]),
    ('kv_sketch_read_table', c_int, [_P, c_int, _P, c_uint64]),
    ('kv_sketch_write_table', c_int, [_P, c_int, _P, c_uint64]),
    ('kv_consume_batch', c_int, [_P, _P, _P, c_uint64, c_int, c_int, c_int, _P, c_int, c_int, POINTER(c_uint64)]),
    ('kv_novel_batch', c_int, [POINTER(_P), c_int, POINTER(_P), c_int, _P, _P, c_uint64, c_int, c_int, c_int,
                               c_int, c_int, c_int64, _P, c_uint64, POINTER(c_uint64), _P, _P]),
    ('kv_hash_kmers', c_int, [c_int, c_int, _P, c_uint64, c_int, _P, _P]),
    ('kv_get_hashes', c_int, [_P, _P, c_uint64, _P]),
    ('kv_add_hashes', c_int, [_P, _P, c_uint64]),
    ('kv_kmer_counts_batch', c_int, [_P, _P, _P, c_uint64, c_int, _P, _P, _P]),
    ('kv_abund_dist_batch', c_int, [_P, _P, _P, _P, c_uint64, c_int, _P]),
    ('kv_sketch_widen', c_int, [_P, _P, POINTER(c_uint64), POINTER(c_int)]),
    ('kv_sketch_narrow', c_int, [_P, _P]),
    ('kv_sketch_merge_peers', c_int, [_P, POINTER(_P), c_int, c_uint64, c_uint64]),
    ('kv_sketch_copy_from_peer', c_int, [_P, _P, c_uint64, c_uint64]),
    ('kv_sketch_allreduce_peers', c_int, [_P, POINTER(_P), c_int, c_uint64, c_uint64]),
    ('kv_sketch_create_shard', c_int, [c_int, c_int, c_int, c_int, POINTER(c_uint64), c_int, c_int, c_int, POINTER(_P)]),
    ('kv_sketch_shard_info', c_int, [_P, POINTER(c_int), POINTER(c_int), POINTER(c_uint64), POINTER(c_uint64)]),
    ('kv_sketch_save_part', c_int, [_P, c_char_p, c_int, c_int, c_uint64]),
    ('kv_sketch_occupancy', c_int, [_P, POINTER(_P), POINTER(c_uint64)]),
    ('kv_unique_batch', c_int, [_P, POINTER(_P), _P, _P, c_uint64, c_int, c_int, c_int, _P, c_int, c_int, POINTER(c_uint64), _P]),
    ('kv_unique_last_batch', c_int, [_P, POINTER(_P), POINTER(c_uint64), _P]),
    ('kv_sketch_set_unique', c_int, [_P, c_uint64]),
    ('kv_sketch_set_unique_dev', c_int, [_P, _P]),
    ('kv_sketch_create_span', c_int, [c_int, c_int, c_int, c_int, POINTER(c_uint64), c_int, c_int, c_int, c_uint64, POINTER(_P),
                                      POINTER(c_int), _P]),
    ('kv_sketch_span_attach', c_int, [_P, c_int, POINTER(c_int), _P]),
    ('kv_sketch_span_ready', c_int, [_P]),
    ('kv_sketch_span_info', c_int, [_P, POINTER(c_int), POINTER(c_int), POINTER(c_uint64), POINTER(c_uint64)]),
    ('kv_consume_batch_span', c_int, [_P, _P, _P, _P, c_uint64, c_int, c_uint64, c_int, c_int, _P, c_int, c_int,
                                      POINTER(c_uint64)]),
    ('kv_peer_sync_create', c_int, [c_int, c_int, c_int, POINTER(_P), _P]),
    ('kv_peer_sync_connect', c_int, [_P, c_int, _P]),
    ('kv_peer_barrier', c_int, [_P]),
    ('kv_peer_sync_status', c_int, [_P]),
    ('kv_peer_sync_destroy', c_int, [_P]),
    ('kv_peer_sync_set_lane', c_int, [_P, c_int]),
    ('kv_merge_fork', c_int, [c_int]),
    ('kv_merge_join', c_int, [c_int]),
    ('kv_hash_batch_dev', c_int, [c_int, c_int, _P, _P, c_uint64, c_int, c_int, c_int, c_int, _P, _P, c_uint64,
                                  POINTER(c_uint64), POINTER(c_uint64)]),
    ('kv_add_hashes_dev', c_int, [_P, _P, _P, c_uint64]),
    ('kv_get_hashes_dev', c_int, [_P, _P, _P, c_uint64, _P]),
    ('kv_novel_from_counts', c_int, [_P, c_int, c_int, POINTER(_P), _P, _P, c_uint64, c_int, c_int, c_int, c_int, c_int,
                                     c_int64, _P, c_uint64, POINTER(c_uint64), _P, _P]),
    ('kv_sketch_ipc_export', c_int, [_P, _P]),
    ('kv_ipc_open', c_int, [c_int, _P, POINTER(_P)]),
    ('kv_ipc_close', c_int, [c_int, _P]),
    ('kv_reader_open', c_int, [c_char_p, POINTER(_P)]),
    ('kv_reader_next', c_int, [_P, c_uint64, POINTER(_P), POINTER(_P), POINTER(c_uint64), POINTER(_P), POINTER(_P),
                               POINTER(_P), POINTER(_P), POINTER(_P)]),
    ('kv_reader_num_reads', c_int, [_P, POINTER(c_uint64)]),
    ('kv_reader_close', c_int, [_P]),
    ('kv_reader_next_batch', c_int, [_P, c_uint64, c_int, POINTER(_P)]),
    ('kv_batch_arrays', c_int, [_P, POINTER(_P), POINTER(_P), POINTER(c_uint64), POINTER(_P), POINTER(_P), POINTER(_P), POINTER(_P),
                                POINTER(_P)]),
    ('kv_batch_release', c_int, [_P]),
    ('kv_synth_reads', c_int, [c_int, POINTER(_P), POINTER(c_uint64), c_int, c_uint64, c_uint64, ctypes.c_uint32,
                               ctypes.c_double, c_uint64, _P, _P]),
    ('kv_stream', c_int, [c_int, POINTER(_P)]),
    ('kv_sync', c_int, [c_int]),
    ('kv_launch_count', c_int, [c_int, POINTER(c_uint64)]),
    ('kv_redo_count', c_int, [c_int, POINTER(c_uint64)]),
    ('kv_profile', c_int, [c_int, c_int, POINTER(ctypes.c_double), POINTER(c_uint64)]),
]


class KvError(RuntimeError):
    """CUDA failure or missing device inside libkvsketch."""


_lib = None


def lib():
    """Load libkvsketch.so once; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise KvError(
                'libkvsketch.so is not built ({}); run `python -c "import __graft_entry__ as g; g.build()"` '
                'or `make -C kevlar_b200/csrc`. kevlar_b200 has no CPU fallback.'.format(LIBPATH))
        handle = ctypes.CDLL(LIBPATH)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.kv_abi_version() != 1:
            raise KvError('libkvsketch.so ABI mismatch')
        _lib = handle
    return _lib


def check(rc):
    """Map a C return code onto the exception type khmer would raise."""
    if rc == KV_OK:
        return
    msg = lib().kv_last_error().decode('utf-8', 'replace')
    if rc == KV_EINVAL:
        raise ValueError(msg)
    if rc == KV_EIO:
        raise OSError(msg)
    if rc == KV_ENOMEM:
        raise MemoryError(msg)
    raise KvError(msg)


def device_count():
    n = c_int()
    check(lib().kv_device_count(byref(n)))
    return n.value


def current_device():
    """Device this process works on: LOCAL_RANK under torchrun, else KV_DEVICE, else 0."""
    for var in ('KV_DEVICE', 'LOCAL_RANK'):
        if os.environ.get(var, '') != '':
            return int(os.environ[var])
    return 0


def primes_below(x, n):
    out = (c_uint64 * n)()
    check(lib().kv_primes_below(int(x), n, out))
    return list(out)


def sync(device=None):
    check(lib().kv_sync(current_device() if device is None else device))


def stream_ptr(device=None):
    p = c_void_p()
    check(lib().kv_stream(current_device() if device is None else device, byref(p)))
    return p.value


def launch_count(device=None):
    n = c_uint64()
    check(lib().kv_launch_count(current_device() if device is None else device, byref(n)))
    return n.value


PROF_CLASSES = ('other', 'hash', 'increment', 'unique', 'novel', 'merge', 'fixup', 'partition')


def profile(enable, device=None):
    """Start (1) / stop (0) / read (2) per-kernel-class timing; returns {class: (ms, launches)}."""
    ms = (ctypes.c_double * len(PROF_CLASSES))()
    n = (c_uint64 * len(PROF_CLASSES))()
    check(lib().kv_profile(current_device() if device is None else device, int(enable), ms, n))
    return {name: (ms[i], n[i]) for i, name in enumerate(PROF_CLASSES)}


def redo_count(device=None):
    """Chunks rolled back and redone exactly after a speculative counter overflow."""
    n = c_uint64()
    check(lib().kv_redo_count(current_device() if device is None else device, byref(n)))
    return n.value


def as_u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def as_u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)
