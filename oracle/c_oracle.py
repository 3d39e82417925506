"""ctypes loader for oracle/libkforacle.so (plain-C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libkforacle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libkforacle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.kfo_vocab_size.argtypes = [ctypes.c_int]
        L.kfo_count_buffer.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p]
        L.kfo_count_buffers_mt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.kfo_count_sparse.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB


def count_buffer(data: bytes, k: int) -> np.ndarray:
    L = lib()
    V = L.kfo_vocab_size(k)
    out = np.zeros(V, dtype=np.uint64)
    tot = ctypes.c_uint64(0)
    rc = L.kfo_count_buffer(data, len(data), k, out.ctypes.data, ctypes.byref(tot))
    if rc != 0:
        raise ValueError("kfo_count_buffer failed: %d" % rc)
    return out


def count_buffers_mt(bufs, k: int, threads: int, want_freq: bool = True):
    """bufs: list of bytes/np.uint8 arrays.  Returns (counts[n,V] u64, freq[n,V] f64 or None, status[n])."""
    L = lib()
    V = L.kfo_vocab_size(k)
    n = len(bufs)
    arrs = [np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b for b in bufs]
    ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = (ctypes.c_size_t * n)(*[a.size for a in arrs])
    counts = np.zeros((n, V), dtype=np.uint64)
    freq = np.zeros((n, V), dtype=np.float64) if want_freq else None
    status = np.zeros(n, dtype=np.int32)
    rc = L.kfo_count_buffers_mt(ptrs, lens, n, k, threads, counts.ctypes.data,
                                freq.ctypes.data if want_freq else None, status.ctypes.data)
    if rc != 0:
        raise ValueError("kfo_count_buffers_mt failed: %d" % rc)
    return counts, freq, status


def count_sparse(data, k: int):
    """Observed canonical k-mers of one file, ascending: (codes u64 [nd], counts u64 [nd], total valid k-mers); k <= 31."""
    L = lib()
    data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    cap = max(len(data), 1)
    codes = np.zeros(cap, dtype=np.uint64)
    counts = np.zeros(cap, dtype=np.uint64)
    nd, tot = ctypes.c_uint64(0), ctypes.c_uint64(0)
    rc = L.kfo_count_sparse(data, len(data), k, codes.ctypes.data, counts.ctypes.data, cap, ctypes.byref(nd), ctypes.byref(tot))
    if rc != 0:
        raise ValueError("kfo_count_sparse failed: %d" % rc)
    return codes[:nd.value].copy(), counts[:nd.value].copy(), int(tot.value)
