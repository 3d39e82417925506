"""Developer experiment: cost per file piece -- same total bytes cut into more and more files (1 contig, no N)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kf2vecfsw_b200 import engine
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kfsynth
from concurrent.futures import ThreadPoolExecutor
engine.init(0)
V = 8192
def run(n, bases, mc=1, nr=0):
    with ThreadPoolExecutor(16) as ex:
        bufs = list(ex.map(lambda i: kfsynth.synth_fasta(1, i, bases, 80, None, mc, nr), range(n)))
    arena = engine.DeviceArena(bufs)
    counts = torch.empty((n, V), dtype=torch.int64, device="cuda")
    ms = []
    for it in range(6):
        engine.count_device(arena, k=7, counts=counts)
        torch.cuda.synchronize()
        ms.append(engine.last_count_kernel_ms())
    best = min(ms[1:])
    print("files=%5d x %9d bases (contigs<=%d, N-runs=%d)  bytes=%.2f GB  kernel %.3f ms  %.2f Tbases/s  %.0f GB/s" % (n, bases, mc, nr, arena.file_bytes/1e9, best, n*bases/best/1e9, arena.file_bytes/best/1e6), flush=True)
    del arena
total = 5_000_000_000
for n in (148, 296, 592, 1000, 2000, 4000):
    run(n, total // n)
run(1000, 5_000_000, 50, 10)
