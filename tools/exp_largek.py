#!/usr/bin/env python
"""One call of the k = 9 (or argv[1]) path on 148 synthetic 5 Mbp genomes: the ncu target for count_fasta_part_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from concurrent.futures import ThreadPoolExecutor
from kf2vecfsw_b200 import engine
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kfsynth
k = int(sys.argv[1]) if len(sys.argv) > 1 else 9
engine.init(0)
with ThreadPoolExecutor(16) as ex:
    fa = list(ex.map(lambda i: kfsynth.synth_fasta(20261018, i, 5_000_000), range(148)))
arena = engine.DeviceArena(fa)
counts = torch.empty((148, engine.vocab_size(k)), dtype=torch.int64, device="cuda")
for it in range(3):
    engine.count_device(arena, k=k, counts=counts)
    torch.cuda.synchronize()
    print("k=%d kernels %.3f ms" % (k, engine.last_count_kernel_ms()))
