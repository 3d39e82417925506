import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from kf2vecfsw_b200 import engine
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kfsynth
from concurrent.futures import ThreadPoolExecutor
engine.init(0)
with ThreadPoolExecutor(8) as ex:
    bufs = list(ex.map(lambda i: kfsynth.synth_fastq(20261018, i, 5_000_000, 1_000_000, 150), range(8)))
arena = engine.DeviceArena(bufs)
counts = torch.empty((8, 8192), dtype=torch.int64, device="cuda")
ms = []
for it in range(6):
    engine.count_device(arena, k=7, counts=counts); torch.cuda.synchronize(); ms.append(engine.last_count_kernel_ms())
print("FASTQ k=7 kernels %.3f ms  %.3f Tbases/s  sum=%d" % (min(ms[1:]), 1.2e9 / min(ms[1:]) / 1e9, int(counts.sum())))
