"""ctypes binding of libkfcount.so (include/kfcount.h) -- the host-side mirror of the C ABI.

No CPU fallback lives here: if the shared library is missing or no sm_100 device is visible, compute
calls raise ``KfError``.  Host-only helpers (vocabulary, .kf formatting and parsing) work without
a GPU because they are plain C++ inside the same library.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

KF_FLAG_PSEUDOCOUNT = 1
KF_FLAG_RAW_CNT = 2
KF_FLAG_FORCE_WALKER = 4
KF_CHUNK = 512
KF_TAIL_PAD = 4096
KF_FLAG_NO_LINEGRID = 8
KF_FLAG_PART_ALL = 16
KF_MAX_K = 12
KF_SPARSE_MIN_K = 6
KF_SPARSE_MAX_K = 31

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_INIT_DEVICE: Optional[int] = None


class KfError(RuntimeError):
    def __init__(self, code: int, what: str = ""):
        self.code = code
        msg = "libkfcount error %d" % code
        try:
            L = _load()
            msg += " (%s)" % L.kf_strerror(code).decode()
            if code == -3:
                msg += ": " + L.kf_last_cuda_error().decode()
        except Exception:  # pragma: no cover
            pass
        if what:
            msg = what + ": " + msg
        super().__init__(msg)


def lib_path() -> str:
    return os.path.join(_HERE, "libkfcount.so")


def _load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            "kf2vecfsw_b200: %s is not built. Run `python -m kf2vecfsw_b200.build` (needs nvcc). "
            "There is no CPU fallback." % path)
    L = ctypes.CDLL(path)
    c_u8pp = ctypes.POINTER(ctypes.c_void_p)
    L.kf_init.argtypes = [ctypes.c_int]
    L.kf_shutdown.argtypes = []
    L.kf_device.argtypes = []
    L.kf_strerror.argtypes = [ctypes.c_int]
    L.kf_strerror.restype = ctypes.c_char_p
    L.kf_last_cuda_error.restype = ctypes.c_char_p
    L.kf_abi_version.restype = ctypes.c_int
    L.kf_vocab_size.argtypes = [ctypes.c_int]
    L.kf_vocab_size.restype = ctypes.c_int64
    L.kf_vocab.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    L.kf_vocab_codes.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    L.kf_count_buffers.argtypes = [c_u8pp, ctypes.POINTER(ctypes.c_size_t), ctypes.c_int, ctypes.c_int,
                                   ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p]
    L.kf_count_files.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.kf_count_device.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.kf_count_windows.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.kf_last_file_status.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p]
    L.kf_last_launch_count.restype = ctypes.c_int
    L.kf_sparse_count.argtypes = [c_u8pp, ctypes.POINTER(ctypes.c_size_t), ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p]
    L.kf_sparse_count_device.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]
    L.kf_sparse_total_entries.restype = ctypes.c_int64
    L.kf_sparse_fetch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.kf_sparse_kmer_matrix.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
    L.kf_sparse_chunk.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p]
    L.kf_set_sm_limit.argtypes = [ctypes.c_int]
    L.kf_set_sm_limit.restype = ctypes.c_int
    L.kf_last_count_kernel_ms.argtypes = [ctypes.POINTER(ctypes.c_float)]
    L.kf_format_row.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_char_p,
                                ctypes.c_size_t]
    L.kf_format_row.restype = ctypes.c_int64
    L.kf_write_kf.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                              ctypes.c_int]
    L.kf_write_kf_rows.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_void_p, ctypes.c_int]
    L.kf_linearise_fasta.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    L.kf_linearise_fasta.restype = ctypes.c_int64
    L.kf_parse_kf.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.kf_parse_kf.restype = ctypes.c_int64
    _LIB = L
    return L


def _check(rc: int, what: str = ""):
    if rc != 0:
        raise KfError(rc, what)


def _flags(pseudocount: bool, raw_cnt: bool, force_walker: bool = False, no_linegrid: bool = False, part_all: bool = False) -> int:
    return (KF_FLAG_PSEUDOCOUNT if pseudocount else 0) | (KF_FLAG_RAW_CNT if raw_cnt else 0) | \
           (KF_FLAG_FORCE_WALKER if force_walker else 0) | (KF_FLAG_NO_LINEGRID if no_linegrid else 0) | \
           (KF_FLAG_PART_ALL if part_all else 0)


# ---- lifecycle -----------------------------------------------------------------------------------
def init(device: Optional[int] = None) -> int:
    """One process per GPU: picks LOCAL_RANK (torchrun) or 0 unless told otherwise."""
    global _INIT_DEVICE
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _INIT_DEVICE == device:
        return device
    _check(_load().kf_init(device), "kf_init(%d)" % device)
    _INIT_DEVICE = device
    return device


def shutdown() -> None:
    """Frees every device and pinned allocation of the library; ``init`` may be called again afterwards."""
    global _INIT_DEVICE
    if _INIT_DEVICE is not None:
        _check(_load().kf_shutdown(), "kf_shutdown")
        _INIT_DEVICE = None


def _require_init():
    if _INIT_DEVICE is None:
        init()


# ---- vocabulary ------------------------------------------------------------------------------------
def vocab_size(k: int) -> int:
    v = _load().kf_vocab_size(k)
    if v < 0:
        raise KfError(int(v), "kf_vocab_size")
    return int(v)


def vocab(k: int) -> List[str]:
    V = vocab_size(k)
    buf = ctypes.create_string_buffer(V * (k + 1))
    _check(_load().kf_vocab(k, buf, len(buf)), "kf_vocab")
    return buf.raw.decode().split("\n")[:V]


def vocab_codes(k: int) -> np.ndarray:
    V = vocab_size(k)
    out = np.zeros(V, dtype=np.uint32)
    _check(_load().kf_vocab_codes(k, out.ctypes.data, V), "kf_vocab_codes")
    return out


# ---- counting: host buffers -------------------------------------------------------------------------
def _as_u8(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        if b.dtype != np.uint8 or not b.flags.c_contiguous:
            raise TypeError("buffers must be contiguous uint8 arrays or bytes")
        return b
    if hasattr(b, "numpy") and hasattr(b, "data_ptr"):  # CPU torch tensor (possibly pinned)
        return b.numpy()
    return np.frombuffer(b, dtype=np.uint8)


def count_buffers(bufs: Sequence, k: int = 7, pseudocount: bool = False, raw_cnt: bool = False,
                  want_counts: bool = True, want_freq: bool = True, force_walker: bool = False,
                  no_linegrid: bool = False, out_counts: Optional[np.ndarray] = None, out_freq: Optional[np.ndarray] = None,
                  part_all: bool = False) -> Tuple[Optional[np.ndarray], Optional[np.ndarray], np.ndarray, np.ndarray]:
    """End-to-end call on host buffers (one per input file).  Returns (counts u64 [n,V] | None,
    freq f64 [n,V] | None, totals u64 [n], status i32 [n])."""
    _require_init()
    L = _load()
    n = len(bufs)
    V = vocab_size(k)
    arrs = [_as_u8(b) for b in bufs]
    ptrs = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data if a.size else None for a in arrs])
    lens = (ctypes.c_size_t * max(n, 1))(*[a.size for a in arrs])
    counts = out_counts if out_counts is not None else (np.empty((n, V), dtype=np.uint64) if want_counts else None)
    freq = out_freq if out_freq is not None else (np.empty((n, V), dtype=np.float64) if want_freq else None)
    totals = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    rc = L.kf_count_buffers(ptrs, lens, n, k, _flags(pseudocount, raw_cnt, force_walker, no_linegrid, part_all),
                            counts.ctypes.data if counts is not None else None,
                            freq.ctypes.data if freq is not None else None, totals.ctypes.data, status.ctypes.data)
    _check(rc, "kf_count_buffers")
    return counts, freq, totals, status


def count_files(paths: Sequence[str], k: int = 7, pseudocount: bool = False, raw_cnt: bool = False):
    _require_init()
    L = _load()
    n = len(paths)
    V = vocab_size(k)
    arr = (ctypes.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])
    counts = np.empty((n, V), dtype=np.uint64)
    freq = np.empty((n, V), dtype=np.float64)
    totals = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    _check(L.kf_count_files(arr, n, k, _flags(pseudocount, raw_cnt), counts.ctypes.data, freq.ctypes.data,
                            totals.ctypes.data, status.ctypes.data), "kf_count_files")
    return counts, freq, totals, status


def count_windows(seq, win_off, win_len, k: int = 7, pseudocount: bool = False, raw_cnt: bool = True,
                  want_freq: bool = False):
    """Chunked-genome mode: one row per window seq[win_off[i] : win_off[i] + win_len[i]] of a linearised sequence.
    Returns (counts u64 [n,V], freq f64 [n,V] | None, totals u64 [n])."""
    _require_init()
    seq = _as_u8(seq)
    win_off = np.ascontiguousarray(win_off, dtype=np.uint64)
    win_len = np.ascontiguousarray(win_len, dtype=np.uint32)
    n = int(win_off.size)
    V = vocab_size(k)
    counts = np.empty((n, V), dtype=np.uint64)
    freq = np.empty((n, V), dtype=np.float64) if want_freq else None
    totals = np.zeros(n, dtype=np.uint64)
    _check(_load().kf_count_windows(seq.ctypes.data if seq.size else None, seq.size, win_off.ctypes.data,
                                    win_len.ctypes.data, n, k, _flags(pseudocount, raw_cnt), counts.ctypes.data,
                                    freq.ctypes.data if freq is not None else None, totals.ctypes.data),
           "kf_count_windows")
    return counts, freq, totals


# ---- sparse counting (large k): observed canonical k-mers only -------------------------------------------
def _sparse_fetch(n: int):
    L = _load()
    total = int(L.kf_sparse_total_entries())
    codes = np.empty(total, dtype=np.uint64)
    counts = np.empty(total, dtype=np.uint32)
    row_off = np.zeros(n + 1, dtype=np.uint64)
    _check(L.kf_sparse_fetch(codes.ctypes.data, counts.ctypes.data, row_off.ctypes.data), "kf_sparse_fetch")
    return codes, counts, row_off


def sparse_fetch_counts(n: int):
    """Counts (u32 [E]) and row offsets (u64 [n + 1]) of the last sparse result; the codes stay on the device."""
    L = _load()
    total = int(L.kf_sparse_total_entries())
    counts = np.empty(total, dtype=np.uint32)
    row_off = np.zeros(n + 1, dtype=np.uint64)
    _check(L.kf_sparse_fetch(None, counts.ctypes.data, row_off.ctypes.data), "kf_sparse_fetch")
    return counts, row_off


def sparse_kmer_matrix(file: int, k: int, n_rows: int, divisor) -> np.ndarray:
    """The FSW feature matrix of file `file` of the last sparse result, expanded on the device (kf_sparse_kmer_matrix):
    [n_rows, k + 1] float32 = k base codes A0 T1 C2 G3 + count / divisor (fp32)."""
    mat = np.empty((n_rows, k + 1), dtype=np.float32)
    if n_rows:
        _check(_load().kf_sparse_kmer_matrix(int(file), ctypes.c_float(float(divisor)), mat.ctypes.data), "kf_sparse_kmer_matrix")
    return mat


def sparse_count(bufs: Sequence, k: int, fetch: bool = True):
    """Sort-and-run-length counting of host buffers (FASTA), KF_SPARSE_MIN_K <= k <= KF_SPARSE_MAX_K: what
    ``jellyfish count -C`` + ``jellyfish dump -c`` list (main.py:135-145), ascending by 2-bit code (A0 C1 G2 T3, first
    base most significant).  Returns (codes u64 [E], counts u32 [E], row_off u64 [n+1], totals u64 [n], status i32 [n]);
    file i's entries are [row_off[i], row_off[i+1]).  With fetch=False the entries stay on the device
    (``sparse_chunks``) and codes/counts are None."""
    _require_init()
    L = _load()
    n = len(bufs)
    arrs = [_as_u8(b) for b in bufs]
    ptrs = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data if a.size else None for a in arrs])
    lens = (ctypes.c_size_t * max(n, 1))(*[a.size for a in arrs])
    nd = np.zeros(n, dtype=np.uint64)
    totals = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    _check(L.kf_sparse_count(ptrs, lens, n, k, nd.ctypes.data, totals.ctypes.data, status.ctypes.data), "kf_sparse_count")
    if not fetch:
        row_off = np.zeros(n + 1, dtype=np.uint64)
        row_off[1:] = np.cumsum(nd)
        return None, None, row_off, totals, status
    codes, counts, row_off = _sparse_fetch(n)
    return codes, counts, row_off, totals, status


def sparse_count_device(arena: "DeviceArena", k: int, stream=None, fetch: bool = True):
    """The same on a device-resident arena.  Synchronises ``stream`` (the output size is only known then)."""
    import torch
    _require_init()
    L = _load()
    if stream is None:
        stream = torch.cuda.current_stream(arena.device)
    handle = stream.cuda_stream if stream.cuda_stream != 0 else 1
    nd = np.zeros(arena.n, dtype=np.uint64)
    totals = np.zeros(arena.n, dtype=np.uint64)
    status = np.zeros(arena.n, dtype=np.int32)
    _check(L.kf_sparse_count_device(ctypes.c_void_p(arena.tensor.data_ptr()), arena.nbytes, arena.offsets.ctypes.data,
                                    arena.lens.ctypes.data, arena.formats.ctypes.data, arena.n, k, nd.ctypes.data,
                                    totals.ctypes.data, status.ctypes.data, ctypes.c_void_p(handle)), "kf_sparse_count_device")
    if not fetch:
        row_off = np.zeros(arena.n + 1, dtype=np.uint64)
        row_off[1:] = np.cumsum(nd)
        return None, None, row_off, totals, status
    codes, counts, row_off = _sparse_fetch(arena.n)
    return codes, counts, row_off, totals, status


def sparse_chunks():
    """Device view of the last sparse result: [(codes_ptr, counts_ptr, n_entries, first_entry, file0, file1), ...]."""
    L = _load()
    out = []
    for i in range(int(L.kf_sparse_chunk_count())):
        pc, pn = ctypes.c_void_p(), ctypes.c_void_p()
        ne, fe = ctypes.c_uint64(), ctypes.c_uint64()
        f0, f1 = ctypes.c_int(), ctypes.c_int()
        _check(L.kf_sparse_chunk(i, ctypes.byref(pc), ctypes.byref(pn), ctypes.byref(ne), ctypes.byref(fe), ctypes.byref(f0),
                                 ctypes.byref(f1)), "kf_sparse_chunk")
        out.append((pc.value, pn.value, int(ne.value), int(fe.value), f0.value, f1.value))
    return out


def sparse_release() -> None:
    _check(_load().kf_sparse_release(), "kf_sparse_release")


# ---- counting: device-resident arena ------------------------------------------------------------------
class DeviceArena:
    """A batch of input files laid out in HBM per the contract of ``kf_count_device``: every file starts
    at a multiple of 512 bytes, gaps are NUL, 1 KiB of NUL follows the last file."""

    def __init__(self, host_buffers: Sequence, device=None, pinned_host=None):
        import torch
        _require_init()
        self.device = torch.device("cuda", _INIT_DEVICE) if device is None else device
        arrs = [_as_u8(b) for b in host_buffers]
        self.n = len(arrs)
        self.lens = np.array([a.size for a in arrs], dtype=np.uint64)
        padded = (self.lens + np.uint64(KF_CHUNK - 1)) // np.uint64(KF_CHUNK) * np.uint64(KF_CHUNK)
        self.offsets = np.zeros(self.n, dtype=np.uint64)
        if self.n > 1:
            self.offsets[1:] = np.cumsum(padded)[:-1]
        self.nbytes = int(padded.sum()) + KF_TAIL_PAD
        self.formats = np.array([a[0] if a.size else 0 for a in arrs], dtype=np.uint8)
        self.tensor = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
        for a, off in zip(arrs, self.offsets):
            # (a file the library rejects -- first byte neither '>' nor '@' -- is left as NUL bytes: a preceding file
            #  that ends without a newline exactly on a 512-byte boundary must not run on into it)
            if a.size and a[0] in (0x3E, 0x40):
                src = torch.from_numpy(a) if a.flags.writeable else torch.frombuffer(memoryview(a), dtype=torch.uint8)
                self.tensor[int(off): int(off) + a.size].copy_(src, non_blocking=True)
        torch.cuda.synchronize(self.device)

    @classmethod
    def empty(cls, lens, first_bytes, device=None):
        """An arena laid out for files of the given lengths (first_bytes: '>' / '@' per file), all NUL; fill with load()."""
        import torch
        _require_init()
        self = cls.__new__(cls)
        self.device = torch.device("cuda", _INIT_DEVICE) if device is None else device
        self.n = len(lens)
        self.lens = np.asarray(lens, dtype=np.uint64).copy()
        padded = (self.lens + np.uint64(KF_CHUNK - 1)) // np.uint64(KF_CHUNK) * np.uint64(KF_CHUNK)
        self.offsets = np.zeros(self.n, dtype=np.uint64)
        if self.n > 1:
            self.offsets[1:] = np.cumsum(padded)[:-1]
        self.nbytes = int(padded.sum()) + KF_TAIL_PAD
        self.formats = np.asarray(first_bytes, dtype=np.uint8).copy()
        self.tensor = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
        return self

    def load(self, i: int, host_array, non_blocking: bool = True) -> None:
        """Copies file i's bytes (uint8 array, pinned for an asynchronous copy) to its place in the arena."""
        import torch
        a = _as_u8(host_array)
        assert a.size == int(self.lens[i])
        off = int(self.offsets[i])
        self.tensor[off: off + a.size].copy_(torch.from_numpy(a), non_blocking=non_blocking)

    @property
    def file_bytes(self) -> int:
        return int(self.lens.sum())


def count_device(arena: DeviceArena, k: int = 7, pseudocount: bool = False, raw_cnt: bool = False,
                 counts=None, freq=None, feat=None, totals=None, force_walker: bool = False, no_linegrid: bool = False,
                 stream=None, part_all: bool = False):
    """Enqueues count + fold/normalise for the arena on ``stream`` (default: torch's current stream).
    Output tensors (torch, on the arena's device) are optional: counts int64/uint64 [n,V], freq float64
    [n,V], feat float32 [n,V] (= fp32(freq*1e4), the matrix the trainers consume), totals int64 [n]."""
    import torch
    _require_init()
    L = _load()
    if stream is None:
        stream = torch.cuda.current_stream(arena.device)
    # torch's default stream has handle 0, which the C ABI reads as "the library's own stream": name the legacy
    # default stream explicitly (cudaStreamLegacy == 0x1) so that the work is ordered with the caller's torch work
    handle = stream.cuda_stream if stream.cuda_stream != 0 else 1
    ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    rc = L.kf_count_device(ctypes.c_void_p(arena.tensor.data_ptr()), arena.nbytes, arena.offsets.ctypes.data,
                           arena.lens.ctypes.data, arena.formats.ctypes.data, arena.n, k,
                           _flags(pseudocount, raw_cnt, force_walker, no_linegrid, part_all), ptr(counts), ptr(freq), ptr(feat),
                           ptr(totals),
                           ctypes.c_void_p(handle))
    _check(rc, "kf_count_device")


def last_file_status(arena: DeviceArena) -> np.ndarray:
    """Per-file status of the last ``count_device`` call on ``arena`` (waits for the device)."""
    status = np.zeros(arena.n, dtype=np.int32)
    _check(_load().kf_last_file_status(ctypes.c_void_p(arena.tensor.data_ptr()), arena.offsets.ctypes.data,
                                       arena.lens.ctypes.data, arena.formats.ctypes.data, arena.n, status.ctypes.data),
           "kf_last_file_status")
    return status


def bind_host_to_gpu(device: int) -> Optional[int]:
    """Pins this process to the CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI function), so
    that the pinned host buffers it allocates afterwards, and the threads that fill them, are local to the GPU's PCIe
    root: with one process per GPU on a two-socket box the host-to-device copies otherwise cross the socket link.
    Returns the NUMA node, or None when the topology cannot be read (nothing is changed then)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use or use == allowed:
            return None
        os.sched_setaffinity(0, use)
        node = int(open(base + "/numa_node").read().strip())
        return node
    except Exception:
        return None


def files_to_kf(in_paths: Sequence[str], out_paths: Sequence[str], samples: Sequence[str], k: int = 7,
                pseudocount: bool = False, raw_cnt: bool = False, threads: int = 0, batch_bytes: int = 0):
    """Files on disk -> .kf files on disk through the pipelined C entry point (reads, GPU, writes overlapped).
    Returns (status i32 [n], totals u64 [n], stage seconds [read wait, gpu, write wait, total])."""
    _require_init()
    L = _load()
    n = len(in_paths)
    assert len(out_paths) == n and len(samples) == n
    enc = lambda xs: (ctypes.c_char_p * max(n, 1))(*[os.fsencode(x) for x in xs])
    status = np.zeros(n, dtype=np.int32)
    totals = np.zeros(n, dtype=np.uint64)
    secs = np.zeros(4, dtype=np.float64)
    L.kf_files_to_kf.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p),
                                 ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p]
    L.kf_files_to_kf.restype = ctypes.c_int
    rc = L.kf_files_to_kf(enc(in_paths), enc(out_paths), enc(samples), n, k, _flags(pseudocount, raw_cnt), int(threads),
                          int(batch_bytes), status.ctypes.data, totals.ctypes.data, secs.ctypes.data)
    _check(rc, "kf_files_to_kf")
    return status, totals, secs


def files_to_device(in_paths: Sequence[str], feat, k: int = 7, pseudocount: bool = False, threads: int = 0, batch_bytes: int = 0):
    """Files on disk -> rows of ``feat`` (torch.float32 CUDA tensor [n, V], contiguous) through the pipelined C entry
    point.  Returns (status i32 [n], totals u64 [n], stage seconds)."""
    _require_init()
    L = _load()
    n = len(in_paths)
    assert feat.is_cuda and feat.is_contiguous() and tuple(feat.shape) == (n, vocab_size(k)) and str(feat.dtype) == "torch.float32"
    paths = (ctypes.c_char_p * max(n, 1))(*[os.fsencode(x) for x in in_paths])
    status = np.zeros(n, dtype=np.int32)
    totals = np.zeros(n, dtype=np.uint64)
    secs = np.zeros(4, dtype=np.float64)
    L.kf_files_to_device.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_int,
                                     ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.kf_files_to_device.restype = ctypes.c_int
    rc = L.kf_files_to_device(paths, n, k, _flags(pseudocount, False), int(threads), int(batch_bytes), ctypes.c_void_p(feat.data_ptr()),
                              status.ctypes.data, totals.ctypes.data, secs.ctypes.data)
    _check(rc, "kf_files_to_device")
    return status, totals, secs


def count_kernel_ms_history(n: int) -> np.ndarray:
    """Durations (ms) of the counting kernels of the last n calls (oldest first, at most 64); waits for them."""
    _require_init()
    L = _load()
    out = np.zeros(max(n, 1), dtype=np.float32)
    L.kf_count_kernel_ms_history.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.kf_count_kernel_ms_history.restype = ctypes.c_int
    w = int(L.kf_count_kernel_ms_history(out.ctypes.data, int(n)))
    if w < 0:
        raise KfError(w, "kf_count_kernel_ms_history")
    return out[:w]


def set_sm_limit(n_sms: int) -> int:
    """Sizes the persistent counting kernels for n_sms SMs (0 = all); returns the SM count in effect."""
    _require_init()
    rc = int(_load().kf_set_sm_limit(int(n_sms)))
    if rc < 0:
        raise KfError(rc, "kf_set_sm_limit")
    return rc


def last_launch_count() -> int:
    return int(_load().kf_last_launch_count())


def last_count_kernel_ms() -> float:
    ms = ctypes.c_float(0)
    _check(_load().kf_last_count_kernel_ms(ctypes.byref(ms)), "kf_last_count_kernel_ms")
    return float(ms.value)


# ---- .kf text -------------------------------------------------------------------------------------------
def format_row(sample: str, row: np.ndarray, int_mode: bool = False) -> str:
    row = np.ascontiguousarray(row, dtype=np.float64)
    s = sample.encode()
    buf = ctypes.create_string_buffer(len(s) + 2 + row.size * 32 + 1)
    n = _load().kf_format_row(s, row.ctypes.data, row.size, 1 if int_mode else 0, buf, len(buf))
    if n < 0:
        raise KfError(int(n), "kf_format_row")
    return buf.raw[:n].decode()


def write_kf(path: str, sample: str, row: np.ndarray, int_mode: bool = False, append: bool = False) -> None:
    row = np.ascontiguousarray(row, dtype=np.float64)
    _check(_load().kf_write_kf(os.fsencode(path), sample.encode(), row.ctypes.data, row.size, 1 if int_mode else 0,
                               1 if append else 0), "kf_write_kf(%s)" % path)


def write_kf_rows(path: str, labels: Sequence[str], rows: np.ndarray, int_modes=None, append: bool = False) -> None:
    """All rows of a chunked-genome .kf file with one open (main.py:895-915)."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    n, V = rows.shape
    lab = b"".join(l.encode() + b"\0" for l in labels)
    im = np.ascontiguousarray(int_modes, dtype=np.uint8) if int_modes is not None else None
    _check(_load().kf_write_kf_rows(os.fsencode(path), lab, rows.ctypes.data, n, V, im.ctypes.data if im is not None else None,
                                    1 if append else 0), "kf_write_kf_rows(%s)" % path)


def linearise_fasta(data, min_len: int):
    """seqtk seq -l 0 + N-run collapse + gap strip + min-length filter (main.py:730-753) in one pass.
    Returns (seq uint8 array, [(header text, seq offset, seq length), ...])."""
    L = _load()
    a = _as_u8(data)
    seq = np.empty(max(a.size, 1), dtype=np.uint8)
    cap = 1024
    while True:
        so = np.zeros(cap, dtype=np.uint64)
        sl = np.zeros(cap, dtype=np.uint64)
        io = np.zeros(cap, dtype=np.uint64)
        il = np.zeros(cap, dtype=np.uint32)
        n = int(L.kf_linearise_fasta(a.ctypes.data if a.size else None, a.size, min_len, seq.ctypes.data, seq.size, so.ctypes.data,
                                     sl.ctypes.data, io.ctypes.data, il.ctypes.data, cap)) if a.size else 0
        if n < 0:
            raise KfError(n, "kf_linearise_fasta")
        if n <= cap:
            break
        cap = n
    recs = [(a[int(io[i]): int(io[i]) + int(il[i])].tobytes().decode("latin-1"), int(so[i]), int(sl[i])) for i in range(n)]
    total = (recs[-1][1] + recs[-1][2]) if recs else 0
    return seq[:total], recs


def parse_kf(text: bytes, V: int, want_rows: bool = True, want_feat: bool = False):
    """Rows of one or more concatenated .kf files.  Returns (labels, rows f64 [n,V] | None, feat f32 [n,V] | None)
    where feat = float32(value * 1e4) (train_classifier_model.py:149,323)."""
    L = _load()
    n = int(L.kf_parse_kf(text, len(text), V, 0, None, None, None, None))
    if n < 0:
        raise KfError(n, "kf_parse_kf")
    rows = np.empty((n, V), dtype=np.float64) if want_rows else None
    feat = np.empty((n, V), dtype=np.float32) if want_feat else None
    off = np.zeros(n, dtype=np.int64)
    ln = np.zeros(n, dtype=np.int32)
    m = int(L.kf_parse_kf(text, len(text), V, n, rows.ctypes.data if rows is not None else None,
                          feat.ctypes.data if feat is not None else None, off.ctypes.data, ln.ctypes.data))
    if m != n:
        raise KfError(m if m < 0 else -5, "kf_parse_kf")
    labels = [text[int(o): int(o) + int(l)].decode() for o, l in zip(off, ln)]
    return labels, rows, feat
