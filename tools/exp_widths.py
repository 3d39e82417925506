#!/usr/bin/env python
"""k = 7 counting throughput by FASTA line width (0 = unwrapped: one line per contig): G x 5 Mbp synthetic genomes resident in
HBM, CUDA events around the whole step; parity of the first genome against the C oracle.  usage: exp_widths.py [G] [width ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from kf2vecfsw_b200 import engine
import kfsynth, c_oracle

G = int(sys.argv[1]) if len(sys.argv) > 1 else 296
widths = [int(x) for x in sys.argv[2:]] or [80, 60, 70, 100, 120, 0]
contigs = int(os.environ.get("KF_CONTIGS", "50"))
nruns = int(os.environ.get("KF_NRUNS", "10"))
engine.init(0)
V = engine.vocab_size(7)
for w in widths:
    with ThreadPoolExecutor(16) as ex:
        fa = list(ex.map(lambda i: kfsynth.synth_fasta(20261018, i, 5_000_000, line_width=(w if w else 10 ** 9), max_contigs=contigs, n_runs=nruns), range(G)))
    arena = engine.DeviceArena(fa)
    counts = torch.empty((G, V), dtype=torch.int64, device="cuda")
    freq = torch.empty((G, V), dtype=torch.float64, device="cuda")
    ms, km = [], []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); engine.count_device(arena, k=7, counts=counts, freq=freq); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1)); km.append(engine.last_count_kernel_ms())
    ref = c_oracle.count_buffer(fa[0].tobytes(), 7)
    ok = bool(np.array_equal(ref, counts[0].cpu().numpy().astype(np.uint64)))
    alg = arena.file_bytes + G * V * 12
    print(json.dumps({"config": "%d x 5 Mbp, line width %s, <= %d contigs, %d N runs" % (G, w or "unwrapped", contigs, nruns), "k": 7, "step_ms": min(ms[1:]), "kernel_ms": min(km[1:]),
                      "gbases_per_s": G * 5e6 / min(ms[1:]) / 1e6, "roofline_frac": alg / min(km[1:]) / 1e6 / 6550.1, "parity_ok": ok}), flush=True)
    if os.environ.get("KF_VL_TIMING"):
        import ctypes
        L = engine._load()
        buf = (ctypes.c_ulonglong * (16 + 640))()
        L.kf_debug_vl_timing(buf, 1)
        engine.count_device(arena, k=7, counts=counts, freq=freq); torch.cuda.synchronize()
        L.kf_debug_vl_timing(buf, 1)
        t = list(buf)
        print(json.dumps({"vl_timing_kclk": {"walk": t[0] / 1e3, "unit_start": t[1] / 1e3, "stage_wait": t[2] / 1e3, "windows": t[3] / 1e3, "total": t[7] / 1e3},
                          "n_walks": t[4], "n_units": t[5], "n_windows": t[6]}), flush=True)
        cta = np.array(t[16:16 + 4 * 148], dtype=np.float64).reshape(148, 4) / 1e3
        o = np.argsort(cta[:, 0])
        print("CTA total kclk: min %.0f median %.0f max %.0f | slowest five (total, pieces, max piece, backscan):" % (cta[o[0], 0], cta[o[74], 0], cta[o[-1], 0]), cta[o[-5:]].round(0).tolist(), flush=True)
    del arena, counts, freq
