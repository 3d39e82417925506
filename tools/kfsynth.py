"""ctypes binding of tools/libkfsynth.so: deterministic synthetic FASTA / FASTQ inputs for bench.py, the tests and the
tools.  Bench / test infrastructure -- the product library does not contain it."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "kf_synth.cpp")
_SO = os.path.join(_HERE, "libkfsynth.so")
_LIB = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-I", _HERE, _SRC, "-o", _SO])
    return _SO


def _load():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.kf_synth_fasta_ex.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        L.kf_synth_fasta_ex.restype = ctypes.c_int64
        L.kf_synth_fastq.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_size_t]
        L.kf_synth_fastq.restype = ctypes.c_int64
        _LIB = L
    return _LIB


def synth_fasta_size(seed: int, genome_id: int, n_bases: int, line_width: int = 80, max_contigs: int = 50,
                     n_runs: int = 10) -> int:
    n = int(_load().kf_synth_fasta_ex(seed, genome_id, n_bases, line_width, max_contigs, n_runs, None, 0))
    if n < 0:
        raise ValueError("kf_synth_fasta_ex: %d" % n)
    return n


def synth_fasta(seed: int, genome_id: int, n_bases: int, line_width: int = 80, out: Optional[np.ndarray] = None,
                max_contigs: int = 50, n_runs: int = 10) -> np.ndarray:
    """line_width = 0: unwrapped (one line per contig)."""
    size = synth_fasta_size(seed, genome_id, n_bases, line_width, max_contigs, n_runs)
    if out is None:
        out = np.empty(size, dtype=np.uint8)
    n = int(_load().kf_synth_fasta_ex(seed, genome_id, n_bases, line_width, max_contigs, n_runs, out.ctypes.data, out.size))
    if n < 0:
        raise ValueError("kf_synth_fasta_ex: %d" % n)
    return out[:n]


def synth_fastq(seed: int, sample_id: int, genome_len: int, n_reads: int, read_len: int = 150) -> np.ndarray:
    L = _load()
    size = int(L.kf_synth_fastq(seed, sample_id, genome_len, n_reads, read_len, None, 0))
    if size < 0:
        raise ValueError("kf_synth_fastq: %d" % size)
    out = np.empty(size, dtype=np.uint8)
    n = int(L.kf_synth_fastq(seed, sample_id, genome_len, n_reads, read_len, out.ctypes.data, out.size))
    if n < 0:
        raise ValueError("kf_synth_fastq: %d" % n)
    return out[:n]
