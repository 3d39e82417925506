"""Drop-in for kf2vec's ``get_chunks(args)`` (reference ``kf2vec/main.py:654-929``): the chunked-genome mode.

The reference shells out, per genome, to ``seqtk seq -l 0`` (linearise, :732), ``awk gsub(/[N|n]+/,"N")``
(collapse N runs, :740), ``seqkit seq -m 10000 -g`` (drop gaps, drop contigs under 10 kbp, :753),
``seqkit split`` (:784), ``seqtk comp`` + ``seqkit sliding`` per contig (:808-824), ``seqkit split`` again
(:837) and then runs ``get_frequencies(raw_cnt=True)`` -- one ``jellyfish count`` + ``dump`` pair per 10-kbp
chunk (:869-881) -- before concatenating the chunk rows (:895-915).  Here the text preparation is a few
``bytes`` operations on the host and all windows of a genome go to the GPU in ONE call
(``kf_count_windows``); the output file, row labels, row order within a contig, thresholds and log lines
are the reference's.  Contig order follows the input file (the reference's ``os.listdir`` order is arbitrary).
"""
from __future__ import annotations

import logging
import math
import os
import sys
import time
from typing import List, Tuple

import numpy as np

from . import engine
from .frequencies import list_inputs, DEFAULT_K

CHUNK_SZ = 10000     # main.py:100
CHUNK_CNT_THR = 5    # main.py:101


def hms(sec_elapsed):   # utils.py:320-328
    h = int(sec_elapsed / (60 * 60))
    m = int((sec_elapsed % (60 * 60)) / 60)
    s = int(sec_elapsed % 60)
    return h, m, s


def window_plan(length: int) -> List[Tuple[int, int]]:
    """1-based inclusive windows of one contig: main.py:813-824 (seqkit sliding emits full windows only)."""
    total_chunks = math.ceil(length / CHUNK_SZ)
    ovrlap = int(math.ceil((total_chunks * CHUNK_SZ - length) / (total_chunks - 1))) if total_chunks != 1 else 0
    step = CHUNK_SZ - ovrlap
    out, s = [], 0
    while s + CHUNK_SZ <= length:
        out.append((s + 1, s + CHUNK_SZ))
        s += step
    return out


def plan_genome(sample: str, data: bytes):
    """Host-side text preparation of one genome (one C++ pass: linearise, collapse N runs, strip gaps, drop contigs
    under 10 kbp; main.py:730-753) and the sliding-window plan (main.py:813-824).
    Returns (sequence buffer uint8, window offsets, window lengths, labels)."""
    seq, recs = engine.linearise_fasta(data, CHUNK_SZ)
    offs, lens, labels = [], [], []
    for header, off, length in recs:
        cid = header.split()[0] if header.split() else ""
        for (a, b) in window_plan(length):
            offs.append(off + a - 1)
            lens.append(b - a + 1)
            # file name seqkit split gives the chunk, minus '.fna' (main.py:895-896)
            labels.append("{}.part_{}.part_{}_sliding__{}-{}".format(sample, cid, cid, a, b))
    return seq, np.array(offs, dtype=np.uint64), np.array(lens, dtype=np.uint32), labels


def chunk_rows(sample: str, data: bytes, k: int = DEFAULT_K, pseudocount: bool = False):
    """(labels, raw counts u64 [n, V]) for one genome; ([], None) if it has fewer than 5 chunks."""
    seq, offs, lens, labels = plan_genome(sample, data)
    if len(labels) < CHUNK_CNT_THR:
        return labels, None
    counts, _, _ = engine.count_windows(seq, offs, lens, k=k)
    return labels, counts


def get_chunks(args) -> None:
    """Reference: kf2vec/main.py:654-929."""
    since = time.time()
    if not os.path.exists(args.input_dir):
        print("No such directory '{}'".format(args.input_dir), file=sys.stderr)
        exit(0)
    if not os.path.exists(args.output_dir):
        print("No such directory '{}'".format(args.output_dir), file=sys.stderr)
        exit(0)

    log = logging.getLogger("kf2vecfsw_b200.get_chunks")
    log.setLevel(logging.INFO)
    log.propagate = False
    for h in list(log.handlers):
        log.removeHandler(h)
    fh = logging.FileHandler(os.path.join(args.output_dir, 'get_chunks_{}.log'.format(
        os.path.basename(os.path.normpath(args.input_dir)))), 'w+')
    sh = logging.StreamHandler()
    for h in (fh, sh):
        h.setFormatter(logging.Formatter('%(message)s'))
        log.addHandler(h)

    def stamp():
        return '{:02d}:{:02d}:{:02d}'.format(*hms(time.time() - since))

    k = getattr(args, 'k', DEFAULT_K)
    pseudocount = bool(getattr(args, 'pseudocount', False))
    log.info('\n==> Making a list of sample names. Time: {}\n'.format(stamp()))
    files_names, samples_names = list_inputs(args.input_dir)
    log.info('\n==> Start processing samples. Time: {}\n'.format(stamp()))

    # Three stages run side by side on successive genomes: a helper thread reads and prepares genome i+1 (one C++ pass),
    # this thread counts the windows of genome i on the GPU, another helper formats and writes the rows of genome i-1.
    # (The library calls release the GIL; the log lines keep the reference's order.)
    from concurrent.futures import ThreadPoolExecutor

    def prepare(fname, sample):
        with open(os.path.join(args.input_dir, fname), "rb") as f:
            data = f.read()
        return plan_genome(sample, data)

    def write_rows(out_path, labels, counts):
        # get_frequencies(raw_cnt=True) rows (main.py:327-357): pandas keeps int64 only when no k-mer is missing
        vals = counts.astype(np.float64)
        if pseudocount:
            vals += 0.5
        int_modes = np.zeros(len(labels), dtype=np.uint8) if pseudocount else (counts > 0).all(axis=1).astype(np.uint8)
        engine.write_kf_rows(out_path, labels, vals, int_modes=int_modes)

    pairs = list(zip(files_names, samples_names))
    with ThreadPoolExecutor(1) as prep_pool, ThreadPoolExecutor(1) as write_pool:
        nxt = prep_pool.submit(prepare, *pairs[0]) if pairs else None
        pending_write = None
        for idx, (fname, sample) in enumerate(pairs):
            log.info('\n==> Start processing. Sample: {}'.format(fname))
            log.info('>>> Formatting to single line. Sample: {}'.format(fname))
            log.info('>>> Replacing stretches of N. Sample: {}'.format(fname))
            log.info('>>> Filtering contigs below threshold {}. Sample: {}'.format(str(CHUNK_SZ), fname))
            seq, offs, lens, labels = nxt.result()
            nxt = prep_pool.submit(prepare, *pairs[idx + 1]) if idx + 1 < len(pairs) else None
            if len(seq) == 0:
                log.info('\n==> Excluded {}. No contigs above threshold length. Time: {}\n'.format(fname, stamp()))
                continue
            log.info('>>> Splitting into contigs. Sample: {}'.format(fname))
            log.info('>>> Getting contig ids. Sample: {}'.format(fname))
            log.info('>>> Computing contig statistics. Sample: {}'.format(fname))
            if len(labels) < CHUNK_CNT_THR:
                log.info('\n==> Excluded {}. {} chunks is too low. {} is required. Time: {}\n'.format(
                    fname, len(labels), CHUNK_CNT_THR, stamp()))
                continue
            log.info('\n==> Done chunk processing for {}. Time: {}\n'.format(fname, stamp()))

            counts, _, _ = engine.count_windows(seq, offs, lens, k=k)
            log.info('\n==> Done computing k-mer frequences for {}. Time: {}\n'.format(fname, stamp()))

            if pending_write is not None:
                pending_write.result()   # (errors of the previous write surface here)
            out_path = os.path.join(args.output_dir, "{}.{}".format(sample, "kf"))
            pending_write = write_pool.submit(write_rows, out_path, labels, counts)
        if pending_write is not None:
            pending_write.result()

    log.info('\n==> Done getting chunks. Time: {}\n'.format(stamp()))
    for h in (fh, sh):
        log.removeHandler(h)
        h.close()
