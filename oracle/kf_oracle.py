"""CPU restatement (NumPy) of kf2vec's k-mer frequency path.  TEST INFRASTRUCTURE ONLY.

This module is the parity oracle: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``kf2vecfsw_b200``) never does.

What it restates (paths relative to the reference checkout):

* ``kf2vec/main.py:250-373``  ``get_frequencies``: file discovery and sample naming (:272-275),
  vocabulary order (:278-296 + ``kf2vec/data/*``), the ``jellyfish count -m k -C`` /
  ``jellyfish dump -c`` pair (:308-319), left-merge on the vocabulary + ``fillna(0)`` (:327-328),
  ``+0.5`` pseudocount (:332-334), fp64 normalisation (:340-342), ``astype(str)`` + join + write
  (:344-357).
* ``kf2vec/main.py:654-929``  ``get_chunks``: linearise, collapse ``[N|n]+`` (:740), drop contigs
  shorter than 10 kbp after gap removal (:753), sliding 10-kbp windows with the computed overlap
  (:813-824), one raw-count row per window (:869-881), row labels (:895-896).

The counting arithmetic itself lives in Jellyfish (external C++, pinned ``kmer-jellyfish=1.1.12`` in
``kf2vec_env.yml:35``, unpinned in ``recipe/meta.yaml:49``), which is not vendored under the
reference tree.  Its published behaviour for ``count -C`` is restated here:

* the file type is sniffed from the first byte (``>`` FASTA, ``@`` FASTQ);
* header lines are skipped, sequence lines are concatenated with ``\\n`` removed, and a break is
  inserted between records, so no k-mer spans two records;
* bases A/C/G/T in either case are coded 0/1/2/3, every other byte (N, IUPAC codes, ``\\r``, ``-``,
  digits ...) resets the window;
* every window of k consecutive coded bases is one k-mer occurrence, binned under
  ``min(kmer, reverse_complement(kmer))`` (2-bit integer compare == lexicographic A<C<G<T);
* FASTQ: sequence lines run until a line starting with ``+``; qualities are skipped by length.

Pinning: the restatement reproduces all seven reproducible golden ``.kf`` files of the reference's
``toy_example`` byte-for-byte and all 358 golden chunk rows count-for-count (see
``tests/test_oracle_golden.py``).  Semantics no reference fixture exercises (lower case, IUPAC, CRLF,
FASTQ, pseudocount, k != 7, empty input) are "parity unpinned": they follow Jellyfish's documented
behaviour only.
"""
from __future__ import annotations

import math
import os
import fnmatch
from typing import Iterable, List, Sequence, Tuple

import numpy as np

BREAK = -1  # window reset symbol

_CODE_LUT = np.full(256, BREAK, dtype=np.int8)
for _ch, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3)):
    _CODE_LUT[ord(_ch)] = _v
    _CODE_LUT[ord(_ch.lower())] = _v

FORMATS = [".fq", ".fastq", ".fa", ".fna", ".fasta"]  # main.py:272


# ----------------------------------------------------------------------------------------------
# vocabulary (main.py:278-296 + kf2vec/data/*): sorted canonical k-mers
# ----------------------------------------------------------------------------------------------
def revcomp_code(x: np.ndarray, k: int) -> np.ndarray:
    """Reverse complement of 2-bit packed k-mers (A0 C1 G2 T3, first base most significant)."""
    x = np.asarray(x, dtype=np.uint64)
    out = np.zeros_like(x)
    for _ in range(k):
        out = (out << np.uint64(2)) | (np.uint64(3) - (x & np.uint64(3)))
        x = x >> np.uint64(2)
    return out


def canonical_codes(k: int) -> np.ndarray:
    """Sorted array of canonical k-mer codes == column order of the .kf row."""
    allk = np.arange(4 ** k, dtype=np.uint64)
    rc = revcomp_code(allk, k)
    return allk[allk <= rc]


def vocab_size(k: int) -> int:
    return (4 ** k + (4 ** (k // 2) if k % 2 == 0 else 0)) // 2


def code_to_kmer(code: int, k: int) -> str:
    return "".join("ACGT"[(int(code) >> (2 * (k - 1 - i))) & 3] for i in range(k))


def vocab(k: int) -> List[str]:
    return [code_to_kmer(c, k) for c in canonical_codes(k)]


# ----------------------------------------------------------------------------------------------
# parsing: file bytes -> symbol stream (codes 0..3, BREAK), newlines removed
# ----------------------------------------------------------------------------------------------
class FormatError(ValueError):
    pass


def _fasta_symbols(buf: np.ndarray) -> np.ndarray:
    n = buf.size
    nl = np.flatnonzero(buf == 10)
    starts = np.concatenate(([0], nl + 1))
    starts = starts[starts < n]
    ends = np.concatenate((nl, [n]))[: starts.size]  # exclusive of the '\n'
    is_hdr = buf[starts] == ord(">")
    line_span = np.diff(np.concatenate((starts, [n])))  # bytes incl. the trailing '\n'
    hdr_byte = np.repeat(is_hdr, line_span)
    keep = (~hdr_byte) & (buf != 10)
    keep[starts[is_hdr]] = True  # the '>' itself stays and decodes to BREAK (record separator)
    return _CODE_LUT[buf[keep]]


def _fastq_symbols(data: bytes) -> np.ndarray:
    """General (multi-line tolerant) FASTQ walk; qualities skipped by length."""
    out: List[np.ndarray] = []
    brk = np.array([BREAK], dtype=np.int8)
    n = len(data)
    pos = 0

    def skip_newlines(p: int) -> int:
        while p < n and data[p] == 10:
            p += 1
        return p

    def ignore_line(p: int) -> int:
        q = data.find(b"\n", p)
        return n if q < 0 else q + 1

    pos = ignore_line(pos)  # first '@' header
    while pos < n:
        nseq = 0
        pos = skip_newlines(pos)
        while pos < n and data[pos] != ord("+"):
            q = data.find(b"\n", pos)
            if q < 0:
                q = n
            out.append(_CODE_LUT[np.frombuffer(data, dtype=np.uint8, count=q - pos, offset=pos)])
            nseq += q - pos
            pos = skip_newlines(q)
        out.append(brk)
        if pos >= n:
            break
        pos = ignore_line(pos)  # '+' line
        nq = 0
        pos = skip_newlines(pos)
        while pos < n and nq < nseq:
            q = data.find(b"\n", pos)
            if q < 0:
                q = n
            nq += q - pos
            pos = skip_newlines(q)
        pos = skip_newlines(pos)
        pos = ignore_line(pos)  # next '@' header
    if not out:
        return np.zeros(0, dtype=np.int8)
    return np.concatenate(out)


def symbols_from_bytes(data: bytes) -> np.ndarray:
    """Jellyfish's view of a sequence file: int8 codes 0..3 and BREAK (-1)."""
    if len(data) == 0:
        raise FormatError("empty file")
    first = data[0]
    if first == ord(">"):
        return _fasta_symbols(np.frombuffer(data, dtype=np.uint8))
    if first == ord("@"):
        return _fastq_symbols(data)
    raise FormatError("unsupported format: first byte %r" % bytes([first]))


# ----------------------------------------------------------------------------------------------
# counting (jellyfish count -C + dump -c + vocabulary merge, main.py:308-328)
# ----------------------------------------------------------------------------------------------
def forward_counts(sym: np.ndarray, k: int) -> np.ndarray:
    """uint64[4^k] counts of every valid forward k-mer window."""
    nb = 4 ** k
    if sym.size < k:
        return np.zeros(nb, dtype=np.uint64)
    s = sym.astype(np.int64)
    m = s.size - k + 1
    code = np.zeros(m, dtype=np.int64)
    bad = np.zeros(m, dtype=bool)
    for j in range(k):
        sl = s[j : j + m]
        bad |= sl < 0
        code = (code << 2) | (sl & 3)
    return np.bincount(code[~bad], minlength=nb).astype(np.uint64)


def fold_canonical(fwd: np.ndarray, k: int) -> np.ndarray:
    """c[m] + c[rc(m)] for canonical m (palindromes counted once), in vocabulary order."""
    can = canonical_codes(k)
    rc = revcomp_code(can, k)
    out = fwd[can.astype(np.int64)].astype(np.uint64)
    nonpal = rc != can
    out[nonpal] += fwd[rc[nonpal].astype(np.int64)]
    return out


def canonical_counts_bytes(data: bytes, k: int) -> np.ndarray:
    return fold_canonical(forward_counts(symbols_from_bytes(data), k), k)


def sparse_counts_bytes(data: bytes, k: int):
    """Observed canonical k-mers only, ascending by code (A0 C1 G2 T3, first base most significant), any k <= 31: what
    `jellyfish count -C` + `jellyfish dump -c` list (kf2vec/main.py:135-145), in sorted instead of hash order.
    Returns (codes uint64 [nd], counts uint64 [nd], total valid k-mers)."""
    assert 1 <= k <= 31
    sym = symbols_from_bytes(data)
    if sym.size < k:
        return np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint64), 0
    s = sym.astype(np.int64)
    m = s.size - k + 1
    fwd = np.zeros(m, dtype=np.uint64)
    rc = np.zeros(m, dtype=np.uint64)
    bad = np.zeros(m, dtype=bool)
    for j in range(k):
        sl = s[j: j + m]
        bad |= sl < 0
        c = (sl & 3).astype(np.uint64)
        fwd = (fwd << np.uint64(2)) | c
        rc |= (np.uint64(3) - c) << np.uint64(2 * j)
    canon = np.minimum(fwd, rc)[~bad]
    codes, counts = np.unique(canon, return_counts=True)
    return codes.astype(np.uint64), counts.astype(np.uint64), int(canon.size)


_FSW_BASE = np.array([0, 2, 3, 1], dtype=np.float32)   # code A0 C1 G2 T3 -> base_map A0 T1 C2 G3 (main.py:118)


def kmer_matrix(data: bytes, k: int):
    """get_kmers' N x (k+1) float32 matrix (kf2vec/main.py:147-172) for one file: k base codes (A0 T1 C2 G3) and
    float32(count) / float32 sum of the counts.  Rows in ascending code order (the reference: Jellyfish hash order;
    the float32 sum therefore agrees to rounding only).  None when no valid k-mer exists (main.py:158-160)."""
    codes, counts, _ = sparse_counts_bytes(data, k)
    if codes.size == 0:
        return None
    mat = np.empty((codes.size, k + 1), dtype=np.float32)
    for j in range(k):
        mat[:, j] = _FSW_BASE[((codes >> np.uint64(2 * (k - 1 - j))) & np.uint64(3)).astype(np.int64)]
    c32 = counts.astype(np.float32)
    mat[:, k] = c32 / np.sum(c32)
    return mat


def canonical_counts_slow(data: bytes, k: int) -> np.ndarray:
    """Pure-Python cross-check that rolls forward and reverse-complement mers like Jellyfish."""
    sym = symbols_from_bytes(data)
    mask = (1 << (2 * k)) - 1
    can = {int(c): i for i, c in enumerate(canonical_codes(k))}
    out = np.zeros(len(can), dtype=np.uint64)
    f = r = 0
    filled = 0
    for c in sym.tolist():
        if c < 0:
            filled = 0
            continue
        f = ((f << 2) | c) & mask
        r = (r >> 2) | ((3 - c) << (2 * (k - 1)))
        filled = min(filled + 1, k)
        if filled == k:
            out[can[min(f, r)]] += 1
    return out


# ----------------------------------------------------------------------------------------------
# pseudocount / normalise / stringify (main.py:332-357)
# ----------------------------------------------------------------------------------------------
def row_values(counts: np.ndarray, pseudocount: bool, raw_cnt: bool) -> Tuple[np.ndarray, bool]:
    """Returns (values, int_mode).  int_mode mirrors the pandas dtype quirk: the merged column is
    int64 only if every vocabulary k-mer was observed (no NaN introduced by the left merge)."""
    int_mode = bool(np.all(counts > 0))
    vals = counts.astype(np.float64)
    if pseudocount:
        vals = vals + 0.5
        int_mode = False
    if not raw_cnt:
        with np.errstate(invalid="ignore", divide="ignore"):
            vals = vals / vals.sum()
        int_mode = False
    return vals, int_mode


def format_value(v: float, int_mode: bool) -> str:
    if int_mode:
        return str(int(v))
    return repr(float(v))


def format_kf_line(sample: str, vals: Sequence[float], int_mode: bool) -> str:
    return sample + "," + ",".join(format_value(v, int_mode) for v in vals) + "\n"


def list_inputs(input_dir: str) -> List[Tuple[str, str]]:
    """(file name, sample name) pairs as main.py:272-275 builds them."""
    files = [f for f in os.listdir(input_dir) if True in (fnmatch.fnmatch(f, "*" + fm) for fm in FORMATS)]
    return [(f, f.rsplit(".f", 1)[0]) for f in files]


def get_frequencies(input_dir: str, output_dir: str, k: int = 7, pseudocount: bool = False,
                    raw_cnt: bool = False) -> List[str]:
    written = []
    for fname, sample in list_inputs(input_dir):
        with open(os.path.join(input_dir, fname), "rb") as fh:
            data = fh.read()
        counts = canonical_counts_bytes(data, k)
        vals, int_mode = row_values(counts, pseudocount, raw_cnt)
        out = os.path.join(output_dir, sample + ".kf")
        with open(out, "w") as fh:
            fh.write(format_kf_line(sample, vals, int_mode))
        written.append(out)
    return written


# ----------------------------------------------------------------------------------------------
# chunked-genome mode (main.py:654-929)
# ----------------------------------------------------------------------------------------------
CHUNK_SZ = 10000      # main.py:100
CHUNK_CNT_THR = 5     # main.py:101


def fasta_records(data: bytes) -> List[Tuple[str, bytes]]:
    """(full header text without '>', linearised sequence) per record (seqtk seq -l 0, main.py:732)."""
    recs: List[Tuple[str, bytes]] = []
    name = None
    parts: List[bytes] = []
    for line in data.split(b"\n"):
        if line.startswith(b">"):
            if name is not None:
                recs.append((name, b"".join(parts)))
            name = line[1:].decode("latin-1")
            parts = []
        elif name is not None:
            parts.append(line.rstrip(b"\r"))
    if name is not None:
        recs.append((name, b"".join(parts)))
    return recs


def collapse_n_runs(seq: bytes) -> bytes:
    """awk gsub(/[N|n]+/,"N") (main.py:740): the class holds 'N', '|' and 'n'."""
    import re
    return re.sub(rb"[N|n]+", b"N", seq)


def strip_gaps(seq: bytes) -> bytes:
    """seqkit seq -g (main.py:753): remove gap characters '-', '.', ' '."""
    return seq.translate(None, b"-. ")


def window_plan(length: int) -> List[Tuple[int, int]]:
    """1-based inclusive (start, end) of the sliding windows of one contig (main.py:813-824)."""
    total_chunks = math.ceil(length / CHUNK_SZ)
    if total_chunks != 1:
        ovrlap = int(math.ceil((total_chunks * CHUNK_SZ - length) / (total_chunks - 1)))
    else:
        ovrlap = 0
    step = CHUNK_SZ - ovrlap
    out = []
    s = 0
    while s + CHUNK_SZ <= length:  # seqkit sliding emits full windows only
        out.append((s + 1, s + CHUNK_SZ))
        s += step
    return out


def chunk_rows(sample: str, data: bytes, k: int = 7) -> List[Tuple[str, np.ndarray]]:
    """Ordered (row label, canonical raw counts) for one genome; [] if the genome is dropped.
    Contig order follows the file (the reference's os.listdir order is arbitrary)."""
    rows: List[Tuple[str, np.ndarray]] = []
    for header, seq in fasta_records(data):
        seq = strip_gaps(collapse_n_runs(seq))
        if len(seq) < CHUNK_SZ:
            continue
        cid = header.split()[0] if header.split() else ""
        for (a, b) in window_plan(len(seq)):
            sub = seq[a - 1 : b]
            sym = _CODE_LUT[np.frombuffer(sub, dtype=np.uint8)]
            cnt = fold_canonical(forward_counts(sym, k), k)
            label = "{}.part_{}.part_{}_sliding__{}-{}".format(sample, cid, cid, a, b)
            rows.append((label, cnt))
    if len(rows) < CHUNK_CNT_THR:
        return []
    return rows
