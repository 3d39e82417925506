import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from kf2vecfsw_b200 import engine
import kfsynth, c_oracle
from fuzzgen import rand_fasta, rand_fasta_grid
engine.init(0)
rng = random.Random(31337)
bufs = []
for i in range(14):
    n = rng.choice([1_000, 30_000, 250_000, 524_288, 700_001, 1_500_000, 3_000_000, 8_000_000])
    b = kfsynth.synth_fasta(99, i, n).tobytes()
    if i % 3 == 0:   # unwrap: one line per contig
        recs = b.split(b">")[1:]
        b = b"".join(b">" + r.split(b"\n", 1)[0] + b"\n" + r.split(b"\n", 1)[1].replace(b"\n", b"") + b"\n" for r in recs)
    bufs.append(b)
bufs += [rand_fasta(rng) for _ in range(5)] + [rand_fasta_grid(rng) for _ in range(5)] + [b"", b">empty\n", b">x\nACGTACGTACGTACGTACGT"]
os.environ["KF_SPARSE_BATCH_BYTES"] = str(6_000_000)   # several sub-batches
for k in (9, 10, 11, 12):
    codes, counts, row_off, totals, status = engine.sparse_count(bufs, k)
    bad = 0
    for i, b in enumerate(bufs):
        if len(b) == 0: continue
        rc, rn, rt = c_oracle.count_sparse(b, k)
        a, e = int(row_off[i]), int(row_off[i + 1])
        ok = rt == int(totals[i]) and np.array_equal(codes[a:e], rc) and np.array_equal(counts[a:e].astype(np.uint64), rn)
        bad += not ok
        if not ok: print("MISMATCH k=%d file %d len %d" % (k, i, len(b)))
    print("k=%d: %d files, %d entries, mismatches %d, chunks %d" % (k, len(bufs), int(row_off[-1]), bad, len(engine.sparse_chunks())), flush=True)
    engine.sparse_release()
