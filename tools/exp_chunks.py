"""Developer experiment: where the time of the chunked-genome mode goes (host prep vs library call)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kf2vecfsw_b200 import engine, chunks
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kfsynth
engine.init(0)
g = [kfsynth.synth_fasta(20261018, i, 5_000_000).tobytes() for i in range(6)]
for i, data in enumerate(g):
    t0 = time.perf_counter()
    seq, offs, lens, labels = chunks.plan_genome("g%d" % i, data)
    t1 = time.perf_counter()
    counts, _, _ = engine.count_windows(np.frombuffer(seq, dtype=np.uint8), offs, lens, k=7)
    t2 = time.perf_counter()
    print("genome %d: %d windows, host plan %.1f ms, kf_count_windows %.1f ms (GPU counting kernels %.3f ms)" % (
        i, len(labels), (t1 - t0) * 1e3, (t2 - t1) * 1e3, engine.last_count_kernel_ms()), flush=True)
