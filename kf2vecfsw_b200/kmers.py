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


def get_kmers(args) -> None:
    """Reference: kf2vec/main.py:112-184."""
    if not os.path.exists(args.output_dir):
        os.makedirs(args.output_dir)
    fasta_files = glob.glob(os.path.join(args.input_dir, "*.fna"))
    for fna_path in fasta_files:
        base_name = os.path.basename(fna_path).replace(".fna", "")
        print(f"--- Processing {base_name} ---")
        counts, _, _, status = engine.count_buffers([np.fromfile(fna_path, dtype=np.uint8)], k=args.k, want_freq=False)
        if status[0] != 0:
            print(f"Error running k-mer counting on {fna_path}: {engine.KfError(int(status[0]))}")
            continue
        if not counts[0].any():
            print(f"Warning: No valid ATCG k-mers found in {base_name}")
            continue
        final_matrix = kmer_matrix(counts[0], args.k)
        output_path = os.path.join(args.output_dir, f"{base_name}_k{args.k}.npy")
        np.save(output_path, final_matrix)
        print(f"Saved: {output_path} (Shape: {final_matrix.shape})")
