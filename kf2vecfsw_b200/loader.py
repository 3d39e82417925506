"""The ``.kf`` loaders of kf2vec's trainers and inference commands, without pandas.

Reference call sites: ``my_read_csv`` (``kf2vec/utils.py:436-437``, one 1-row DataFrame per file through
``mp.Pool``), ``train_classifier_model.py:144-150`` / ``train_model_set.py:285-292`` (concat, ``* 1e4``),
``classify.py:102-114`` (blocks of 4000 files), ``query.py:148-166`` (``cat`` of the files, one ``read_csv``),
and the chunk readers ``utils.py:402-431`` (uint16 rows; capped at 255 for uint8).  All of them end in the same
tensor: ``float32(float64(value) * 1e4)`` of shape ``[rows, V]`` -- built here by one C++ pass over the text
(``kf_parse_kf``), or taken straight from HBM with ``frequency_matrix`` when the .kf round trip is not needed.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np

from . import engine

FEATURES_SCALER = 1e4   # train_classifier_model.py:69


def read_kf(path: str, V: int = None) -> Tuple[List[str], np.ndarray]:
    """``my_read_csv`` (utils.py:436-437): (row labels, float64 [rows, V])."""
    with open(path, "rb") as f:
        text = f.read()
    if V is None:
        first = text.split(b"\n", 1)[0]
        V = first.count(b",")
    labels, rows, _ = engine.parse_kf(text, V)
    return labels, rows


def load_kf_files(paths: Sequence[str], V: int = None, device=None):
    """The trainers' feature matrix: (labels, torch.float32 [N, V]) = float32(value * 1e4), files concatenated in
    the given order (train_classifier_model.py:144-150; query.py:148-166)."""
    import torch
    text = b"".join(open(p, "rb").read() for p in paths)
    if V is None:
        V = text.split(b"\n", 1)[0].count(b",") if text else 0
    labels, _, feat = engine.parse_kf(text, V, want_rows=False, want_feat=True)
    t = torch.from_numpy(feat)
    return labels, (t.to(device) if device is not None else t)


def load_kf_dir(input_dir: str, V: int = None, device=None):
    """glob('<dir>/*.kf') as train_classifier does (main.py:379), sorted for determinism."""
    paths = sorted(os.path.join(input_dir, f) for f in os.listdir(input_dir) if f.endswith(".kf"))
    return load_kf_files(paths, V=V, device=device)


def read_chunk_kf(path: str, V: int = None, cap_uint8: bool = False) -> Tuple[List[str], np.ndarray]:
    """Chunk rows: utils.py:402-405 (uint16) or :416-431 (``min(int(float(v)), 255)`` -> uint8)."""
    labels, rows = read_kf(path, V)
    if cap_uint8:
        return labels, np.minimum(rows.astype(np.int64), 255).astype(np.uint8)
    return labels, rows.astype(np.uint16)
