#!/usr/bin/env python
"""Sparse (sort-and-run-length) path timing: G x 5 Mbp synthetic genomes resident in HBM, k from argv; parity of the first
genomes against the C oracle.  usage: exp_sparse.py [G] [k ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from kf2vecfsw_b200 import engine
import kfsynth, c_oracle

G = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ks = [int(x) for x in sys.argv[2:]] or [12]
engine.init(0)
with ThreadPoolExecutor(16) as ex:
    fa = list(ex.map(lambda i: kfsynth.synth_fasta(20261018, i, 5_000_000), range(G)))
arena = engine.DeviceArena(fa)
for k in ks:
    ms = []
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, row_off, totals, status = engine.sparse_count_device(arena, k, fetch=False)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - t0) * 1e3)
    codes, counts, row_off, totals, status = engine.sparse_count_device(arena, k)
    ok = True
    for i in range(min(2, G)):
        rc, rn, rt = c_oracle.count_sparse(fa[i].tobytes(), k)
        a, e = int(row_off[i]), int(row_off[i + 1])
        ok = ok and rt == int(totals[i]) and np.array_equal(codes[a:e], rc) and np.array_equal(counts[a:e].astype(np.uint64), rn)
    print(json.dumps({"config": "sparse sort-and-run-length, %d x 5 Mbp genomes" % G, "k": k, "ms": min(ms[1:]), "ms_all": ms,
                      "gbases_per_s": G * 5e6 / min(ms[1:]) / 1e6, "entries": int(row_off[-1]), "launches": engine.last_launch_count(),
                      "parity_ok": bool(ok)}), flush=True)
    engine.sparse_release()
