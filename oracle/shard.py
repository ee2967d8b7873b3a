"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes loader for shard_oracle.c (reference util/data.py:11-30)."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build():
    so = _HERE / 'libshard_oracle.so'
    src = _HERE / 'shard_oracle.c'
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(['gcc', '-O2', '-shared', '-fPIC', '-o', str(so), str(src)])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(str(build()))
        L.oracle_shard_indices.restype = ctypes.c_int64
        L.oracle_shard_indices.argtypes = [ctypes.c_int64] * 3 + [ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                                                  ctypes.c_void_p]
        L.oracle_permutation.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p]
        L.oracle_mt_raw.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int]
        _LIB = L
    return _LIB


def shard_indices(num_elements, rank, world_size, shuffle=False, even_shards=True, seed=0):
    """reference util/data.py:11-30, integer-exact."""
    out = np.empty(num_elements // max(world_size, 1) + 2, dtype=np.int64)
    k = _lib().oracle_shard_indices(num_elements, rank, world_size, int(shuffle), int(even_shards), seed,
                                    out.ctypes.data)
    return out[:k].tolist()


def permutation(n, shuffle, seed):
    out = np.empty(max(n, 1), dtype=np.int64)
    _lib().oracle_permutation(n, int(shuffle), seed, out.ctypes.data)
    return out[:n]


def mt_raw(seed, n):
    out = np.empty(n, dtype=np.uint32)
    _lib().oracle_mt_raw(seed, out.ctypes.data, n)
    return out
