#!/usr/bin/env python
"""Small inputs through every counting kernel, for compute-sanitizer (memcheck / racecheck / synccheck):
k = 7 line kernel + generic + FASTQ, k = 8 (one text pass), k = 9 and 10 (text pass + stream passes), chunk windows."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
from kf2vecfsw_b200 import engine
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kfsynth
from fuzzgen import rand_fasta, rand_fasta_grid, rand_fastq
import kf_oracle as o
engine.init(0)
rng = random.Random(5)
bufs = [kfsynth.synth_fasta(1, 0, 600_000).tobytes(), rand_fasta(rng), rand_fasta_grid(rng), rand_fastq(rng),
        kfsynth.synth_fastq(1, 0, 100_000, 2_000, 150).tobytes(), kfsynth.synth_fasta(1, 1, 300_000).tobytes()]
for k in (7, 8, 9, 10):
    for kw in ({}, {"part_all": True}):
        counts, freq, totals, status = engine.count_buffers(bufs, k=k, **kw)
        for i, b in enumerate(bufs):
            assert np.array_equal(counts[i], o.canonical_counts_bytes(bytes(b), k)), (k, kw, i)
    print("k=%d ok" % k, flush=True)
seq = np.frombuffer(("ACGTTGCA" * 5000).encode(), dtype=np.uint8)
offs = np.arange(0, 30000, 2500, dtype=np.uint64)
lens = np.full(len(offs), 10000, dtype=np.uint32)
engine.count_windows(seq, offs, lens, k=7)
print("windows ok")
