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
# sparse path: 4,096-bucket pipeline (k = 7, 15), 16-bit-bucket pipeline (k = 9, 12: write-combined partition, packed counters,
# a poly-A run that forces the exact redo), the device-side k-mer matrix
import c_oracle
fa = [b for b in bufs if bytes(b[:1]) == b">"]
seq = "A" * 70001 + "ACGTTGCAAGGCTTAACCGGTTAA" * 40
fa.append((">poly\n" + "\n".join(seq[j:j + 80] for j in range(0, len(seq), 80)) + "\n").encode())
for k in (7, 9, 12, 15):
    codes, counts, row_off, totals, status = engine.sparse_count(fa, k)
    for i, b in enumerate(fa):
        rc, rn, rt = c_oracle.count_sparse(bytes(b), k)
        a, e = int(row_off[i]), int(row_off[i + 1])
        assert rt == int(totals[i]) and np.array_equal(codes[a:e], rc) and np.array_equal(counts[a:e].astype(np.uint64), rn), (k, i)
    m = engine.sparse_kmer_matrix(0, k, int(row_off[1] - row_off[0]), np.sum(counts[:int(row_off[1])].astype(np.float32)))
    assert m.shape[1] == k + 1 and abs(float(m[:, k].sum()) - 1.0) < 1e-3
    engine.sparse_release()
    print("sparse k=%d ok" % k, flush=True)
