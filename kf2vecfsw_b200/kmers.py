"""Drop-in for the FSW fork's ``get_kmers(args)`` (reference ``kf2vec/main.py:112-184``): per ``*.fna`` file an
``N x (k+1)`` float32 ``.npy`` holding every OBSERVED canonical k-mer as k base codes (A0 T1 C2 G3, main.py:118)
plus its normalised count in the last column -- the set the FSW embedding consumes (train_model_set.py:192-204).

The reference gets the k-mers from ``jellyfish count -C`` + ``jellyfish dump -c -t`` (:135-145) and therefore lists
them in Jellyfish's hash order; rows here are in sorted canonical order (the set embedding is permutation
invariant, so only set equality is defined).  Normalisation is the reference's float32 arithmetic (:165-169).
"""
from __future__ import annotations

import glob
import os

import numpy as np

from . import engine

_STD_TO_FSW = np.array([0, 2, 3, 1], dtype=np.float32)   # vocabulary code A0 C1 G2 T3 -> A0 T1 C2 G3


def kmer_matrix(counts: np.ndarray, k: int) -> np.ndarray:
    """counts: canonical counts in vocabulary order -> N x (k+1) float32 (observed k-mers only)."""
    codes = engine.vocab_codes(k)
    nz = np.flatnonzero(counts)
    c = codes[nz].astype(np.uint32)
    mat = np.empty((nz.size, k + 1), dtype=np.float32)
    for j in range(k):
        mat[:, j] = _STD_TO_FSW[(c >> np.uint32(2 * (k - 1 - j))) & np.uint32(3)]
    cnt = counts[nz].astype(np.float32)
    mat[:, k] = cnt / np.sum(cnt)
    return mat


def kmer_matrix_sparse(codes: np.ndarray, counts: np.ndarray, k: int) -> np.ndarray:
    """Observed canonical k-mers (2-bit codes A0 C1 G2 T3, first base most significant) + counts -> N x (k+1) float32."""
    mat = np.empty((codes.size, k + 1), dtype=np.float32)
    c = codes.astype(np.uint64)
    for j in range(k):
        mat[:, j] = _STD_TO_FSW[((c >> np.uint64(2 * (k - 1 - j))) & np.uint64(3)).astype(np.intp)]
    cnt = counts.astype(np.float32)
    mat[:, k] = cnt / np.sum(cnt)
    return mat


_BATCH_BYTES = 1 << 30   # files are counted in batches of about this many input bytes (one call each)
_DENSE_MAX_K = 10        # up to here the dense canonical row is the faster way to the observed k-mers (k = 7, the default:
                         # 2.7 Tbases/s against 0.03 on the sparse path; k = 10: 0.14 against 0.13; a row is 4 MB at k = 10)


def kmer_matrices(paths, k: int):
    """Yields (path, matrix | None, status) per file.  k >= 11 goes through the sparse sort-and-run-length path
    (``kf_sparse_count``: any k up to 31, the reference's -k range, main.py:81-82), smaller k through the dense rows."""
    paths = list(paths)
    i = 0
    while i < len(paths):
        j, nbytes = i, 0
        while j < len(paths) and (j == i or nbytes + os.path.getsize(paths[j]) <= _BATCH_BYTES):
            nbytes += os.path.getsize(paths[j])
            j += 1
        bufs = [np.fromfile(p, dtype=np.uint8) for p in paths[i:j]]
        if k > _DENSE_MAX_K:
            # the entries stay on the device; a file's matrix is expanded there (kf_sparse_kmer_matrix) and copied out once.
            # The divisor is the reference's float32 sum of the float32 counts (main.py:165-169), taken with NumPy.
            _, _, _, _, status = engine.sparse_count(bufs, k, fetch=False)
            counts, row_off = engine.sparse_fetch_counts(len(bufs))
            for t, p in enumerate(paths[i:j]):
                a, b = int(row_off[t]), int(row_off[t + 1])
                mat = engine.sparse_kmer_matrix(t, k, b - a, np.sum(counts[a:b].astype(np.float32))) if b > a else None
                yield p, mat, int(status[t])
            engine.sparse_release()
        else:
            cnt, _, _, status = engine.count_buffers(bufs, k=k, want_freq=False)
            for t, p in enumerate(paths[i:j]):
                yield p, (kmer_matrix(cnt[t], k) if cnt[t].any() else None), int(status[t])
        i = j


def get_kmers(args) -> None:
    """Reference: kf2vec/main.py:112-184 (same files, prints and .npy layout; rows in sorted instead of hash order)."""
    if not os.path.exists(args.output_dir):
        os.makedirs(args.output_dir)
    fasta_files = glob.glob(os.path.join(args.input_dir, "*.fna"))
    for fna_path, final_matrix, status in kmer_matrices(fasta_files, args.k):
        base_name = os.path.basename(fna_path).replace(".fna", "")
        print(f"--- Processing {base_name} ---")
        if status not in (0, -9):   # (an empty file: Jellyfish reports no k-mers, the reference warns and goes on)
            print(f"Error running k-mer counting on {fna_path}: {engine.KfError(status)}")
            continue
        if final_matrix is None:
            print(f"Warning: No valid ATCG k-mers found in {base_name}")
            continue
        output_path = os.path.join(args.output_dir, f"{base_name}_k{args.k}.npy")
        np.save(output_path, final_matrix)
        print(f"Saved: {output_path} (Shape: {final_matrix.shape})")
