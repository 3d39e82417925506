"""Developer experiment: FASTQ kernel throughput (config 4 shape: 150-bp reads, N-containing) and chunked mode."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kf2vecfsw_b200 import engine
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kfsynth
from concurrent.futures import ThreadPoolExecutor
engine.init(0)
n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
with ThreadPoolExecutor(8) as ex:
    bufs = list(ex.map(lambda i: kfsynth.synth_fastq(20261018, i, 5_000_000, n_reads, 150), range(n_samples)))
arena = engine.DeviceArena(bufs)
for k in (7, 9):
    V = engine.vocab_size(k)
    counts = torch.empty((n_samples, V), dtype=torch.int64, device="cuda")
    ms = []
    for it in range(6):
        engine.count_device(arena, k=k, counts=counts)
        torch.cuda.synchronize()
        ms.append(engine.last_count_kernel_ms())
    best = min(ms[1:])
    bases = n_samples * n_reads * 150
    print("FASTQ k=%d  %d samples x %d reads  bytes=%.2f GB  kernels %.3f ms  %.3f Tbases/s  %.0f GB/s (%.1f%% of 6550)  status=%s" % (
        k, n_samples, n_reads, arena.file_bytes / 1e9, best, bases / best / 1e9, arena.file_bytes / best / 1e6,
        arena.file_bytes / best / 1e6 / 65.5, engine.last_file_status(arena).tolist()))
del arena
# chunked-genome mode: all 10-kbp windows of synthetic genomes
from kf2vecfsw_b200 import chunks
g = [kfsynth.synth_fasta(20261018, i, 5_000_000).tobytes() for i in range(4)]
t0 = time.perf_counter()
nwin = 0
for i, data in enumerate(g):
    labels, counts = chunks.chunk_rows("g%d" % i, data)
    nwin += len(labels)
dt = time.perf_counter() - t0
print("chunked mode: 4 genomes x 5 Mbp, %d windows, %.1f ms per genome end to end (host prep + H2D + kernels + D2H)" % (nwin, dt / 4 * 1e3))
