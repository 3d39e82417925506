#!/usr/bin/env python
"""Extended GPU fuzz of the dense path through kf_count_buffers (its sub-batch pipeline cut at random sizes): random batches
of FASTA (ragged, wrapped at fixed widths, long lines) and FASTQ (4-line and multi-line) files, k drawn from 3..10, against
the NumPy / C oracles.  usage: fuzz_dense_gpu.py [batches] [seed]"""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
from kf2vecfsw_b200 import engine
import c_oracle, kfsynth
from fuzzgen import rand_fasta, rand_fasta_grid, rand_fastq, rand_fastq_multiline
engine.init(0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
bad = 0
for it in range(nb):
    bufs = []
    for _ in range(rng.randint(1, 14)):
        r = rng.random()
        if r < 0.30: b = rand_fasta(rng)
        elif r < 0.55: b = rand_fasta_grid(rng)
        elif r < 0.75: b = rand_fastq(rng)
        elif r < 0.80: b = rand_fastq_multiline(rng)
        elif r < 0.90: b = kfsynth.synth_fasta(rng.randint(0, 10 ** 6), rng.randint(0, 99), rng.choice([5_000, 80_000, 600_000, 1_300_000])).tobytes()
        elif r < 0.95:
            s = "".join(rng.choice("ACGT") for _ in range(rng.randint(50_000, 200_000)))
            b = (">long\n" + s + "\n>two\n" + s[:7777] + "\n").encode()
        else: b = rng.choice([b"", b">h\n", b"junk\n", b"@r\nACGT\n+\nIIII\n"])
        bufs.append(b)
    k = rng.choice([3, 5, 7, 7, 7, 8, 9, 10])
    os.environ["KF_SUB_BATCH_BYTES"] = str(rng.choice([1_000, 50_000, 700_000, 1 << 29]))
    counts, freq, totals, status = engine.count_buffers(bufs, k=k)
    ref, _, _ = c_oracle.count_buffers_mt([np.frombuffer(b, dtype=np.uint8) for b in bufs], k, 8, want_freq=False)
    for i, b in enumerate(bufs):
        if status[i] != 0:
            assert len(b) == 0 or b[:1] not in (b">", b"@"), (it, i, int(status[i]))
            continue
        if not np.array_equal(counts[i], ref[i]):
            bad += 1
            print("MISMATCH batch %d file %d k=%d len %d first %r" % (it, i, k, len(b), b[:20]), flush=True)
print("batches %d, mismatches %d" % (nb, bad))
sys.exit(1 if bad else 0)
