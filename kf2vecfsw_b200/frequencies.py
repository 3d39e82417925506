"""Drop-in for kf2vec's ``get_frequencies(args)`` (reference ``kf2vec/main.py:250-373``).

Same argparse ``Namespace`` in, same ``<output_dir>/<sample>.kf`` files out, same prints, same sample
naming, same column order / pseudocount / normalisation / text format -- but the per-file
``jellyfish count`` + ``jellyfish dump`` subprocesses and the pandas merge are replaced by one batched call
into ``libkfcount.so`` (CUDA, sm_100a).  ``build_library`` (main.py:573), ``process_query_data`` (main.py:629)
and ``get_chunks`` (main.py:869-881) call this function with their own ``args``; both wrapper parsers omit
``raw_cnt`` (main.py:1253-1353), so optional attributes are read with ``getattr``.
"""
from __future__ import annotations

import fnmatch
import os
import sys
from typing import List, Optional, Tuple

import numpy as np

from . import engine

FORMATS = ['.fq', '.fastq', '.fa', '.fna', '.fasta']  # main.py:272
DEFAULT_K = 7                                          # main.py:80
BATCH_BYTES = 256 << 20                                # host bytes staged per pipeline batch (two pinned slabs)


def list_inputs(input_dir: str) -> Tuple[List[str], List[str]]:
    """main.py:272-275 -- os.listdir order, fnmatch on the five suffixes, sample = name.rsplit('.f', 1)[0]."""
    files_names = [f for f in os.listdir(input_dir)
                   if True in (fnmatch.fnmatch(f, '*' + form) for form in FORMATS)]
    samples_names = [f.rsplit('.f', 1)[0] for f in files_names]
    return files_names, samples_names


def get_frequencies(args) -> None:
    """Reference: kf2vec/main.py:250-373."""
    print('\n==> Starting k-mer counting for {}\n'.format(args.input_dir))

    if not os.path.exists(args.input_dir):                       # main.py:255-259
        print("No such directory '{}'".format(args.input_dir), file=sys.stderr)
        exit(0)
    if not os.path.exists(args.output_dir):                      # main.py:262-266
        print("No such directory '{}'".format(args.output_dir), file=sys.stderr)
        exit(0)

    k = getattr(args, 'k', DEFAULT_K)
    pseudocount = bool(getattr(args, 'pseudocount', False))
    raw_cnt = bool(getattr(args, 'raw_cnt', False))

    files_names, samples_names = list_inputs(args.input_dir)
    paths = [os.path.join(args.input_dir, f) for f in files_names]
    V = engine.vocab_size(k)

    # One pipelined library call: host threads read the next batch of files into pinned memory while the GPU counts the
    # current one and the rows of the previous one are formatted and written (kf_files_to_kf).  args.p -- jellyfish's
    # thread count in the reference (main.py:309) -- is the number of those host threads.
    outs = [os.path.join(args.output_dir, "{}.{}".format(s, "kf")) for s in samples_names]
    threads = int(getattr(args, 'p', 0) or 0)
    status, totals, _ = engine.files_to_kf(paths, outs, [str(s) for s in samples_names], k=k, pseudocount=pseudocount,
                                           raw_cnt=raw_cnt, threads=threads, batch_bytes=BATCH_BYTES)
    # The reference prints these lines while it works through the files (main.py:333,341); the files are counted in one
    # pipelined call here, so the same lines, in the same order, follow it.
    for i in range(len(paths)):
        if status[i] != 0:
            # The reference ignores jellyfish's exit code and then dies with IndexError at main.py:315;
            # fail with the real cause instead of writing a bogus row.
            raise engine.KfError(int(status[i]), "k-mer counting failed for {}".format(files_names[i]))
        if pseudocount:
            print('>>> Adding pseudocounts. Sample: {}'.format(files_names[i]))       # main.py:333
        if not raw_cnt:
            print('>>> Normalizing. Sample: {}'.format(files_names[i]))               # main.py:341

    print('\n==> Done processing {}'.format(args.input_dir))


def frequency_matrix(input_dir: str, k: int = DEFAULT_K, pseudocount: bool = False, device=None):
    """Fast path for the trainers: the [N, V] float32 feature matrix fp32(freq * 1e4) straight from HBM,
    skipping the .kf text round trip of train_classifier_model.py:144-150 / utils.py:436-437.
    Returns (sample names, torch.float32 CUDA tensor)."""
    import torch
    files_names, samples_names = list_inputs(input_dir)
    paths = [os.path.join(input_dir, f) for f in files_names]
    V = engine.vocab_size(k)
    dev_index = engine.init() if device is None else engine.init(torch.device(device).index or 0)
    feat = torch.empty((len(paths), V), dtype=torch.float32, device=torch.device("cuda", dev_index))
    if paths:
        # pipelined reads (host threads, pinned slabs) + GPU stage; the rows never leave the device
        status, _, _ = engine.files_to_device(paths, feat, k=k, pseudocount=pseudocount, batch_bytes=BATCH_BYTES)
        for i in range(len(paths)):
            if status[i] != 0:
                raise engine.KfError(int(status[i]), "k-mer counting failed for {}".format(files_names[i]))
        torch.cuda.synchronize(feat.device)
    return samples_names, feat
