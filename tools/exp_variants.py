"""Developer experiment: line-grid kernel time vs data shape (contigs / N-runs / number of files)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kf2vecfsw_b200 import engine
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kfsynth
from concurrent.futures import ThreadPoolExecutor
engine.init(0)
V = 8192
def run(tag, n, bases, mc, nr):
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
    print("%-40s files=%5d bytes=%.2f GB  kernel %.3f ms  %.2f Tbases/s  %.0f GB/s" % (tag, n, arena.file_bytes/1e9, best, n*bases/best/1e9, arena.file_bytes/best/1e6))
    del arena
run("1000 x 5M, <=50 contigs, 10 N-runs", 1000, 5_000_000, 50, 10)
run("1000 x 5M, 1 contig, 10 N-runs", 1000, 5_000_000, 1, 10)
run("1000 x 5M, <=50 contigs, 0 N-runs", 1000, 5_000_000, 50, 0)
run("1000 x 5M, 1 contig, 0 N-runs", 1000, 5_000_000, 1, 0)
run("100 x 50M, 1 contig, 0 N-runs", 100, 50_000_000, 1, 0)
run("148 x 34M, 1 contig, 0 N-runs", 148, 34_000_000, 1, 0)
