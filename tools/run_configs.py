#!/usr/bin/env python
"""BASELINE.json configs[3] (FASTQ query reads) and configs[4] (large-k sweep, chunked-genome mode) on one B200, each
with the CPU oracle timed beside it on a bounded sample.  One JSON line per measurement (kept under profiles/).
Inputs resident in HBM, CUDA events around the counting kernels (kf_last_count_kernel_ms) and around the whole step
(counting + fold/normalise)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from kf2vecfsw_b200 import engine, chunks
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kfsynth
import c_oracle

engine.init(0)
PEAK = 6550.1
SEED = 20261018
threads = len(os.sched_getaffinity(0))


def timed(arena, k, n, V, reps=6):
    counts = torch.empty((n, V), dtype=torch.int64, device="cuda")
    freq = torch.empty((n, V), dtype=torch.float64, device="cuda")
    step, kern = [], []
    for it in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.count_device(arena, k=k, counts=counts, freq=freq)
        e1.record()
        torch.cuda.synchronize()
        step.append(e0.elapsed_time(e1)); kern.append(engine.last_count_kernel_ms())
    return min(step[1:]), min(kern[1:]), counts


def cpu(bufs, k):
    t0 = time.perf_counter()
    ref, _, _ = c_oracle.count_buffers_mt(bufs, k, threads, want_freq=False)
    return time.perf_counter() - t0, ref


def emit(**kw):
    print(json.dumps(kw), flush=True)


# ---- config 4: FASTQ query reads (150 bp, 30x of 5 Mbp, N-containing), 8 samples ----
n_samples, n_reads = 8, 1_000_000
with ThreadPoolExecutor(8) as ex:
    fq = list(ex.map(lambda i: kfsynth.synth_fastq(SEED, i, 5_000_000, n_reads, 150), range(n_samples)))
arena = engine.DeviceArena(fq)
bases = n_samples * n_reads * 150
for k in (7,):
    V = engine.vocab_size(k)
    step, kern, counts = timed(arena, k, n_samples, V)
    dt, ref = cpu(fq[:2], k)
    ok = bool(np.array_equal(ref, counts[:2].cpu().numpy().astype(np.uint64)))
    emit(config="configs[3] FASTQ reads 150 bp x 1e6 x 8 samples", k=k, gbases_per_s=bases / step / 1e6, kernel_ms=kern, step_ms=step,
         file_bytes=arena.file_bytes, roofline_frac=(arena.file_bytes / kern / 1e6) / PEAK, bytes_per_base=arena.file_bytes / bases,
         cpu_gbases_per_s=2 * n_reads * 150 / dt / 1e9, cpu_threads=threads, cpu_sample="2 samples, oracle/kf_oracle.c", parity_ok=ok,
         status=engine.last_file_status(arena).tolist())
del arena

# ---- config 5: large-k sweep on 5 Mbp genomes (whole-genome mode) ----
G = 592   # 4 x 148: (file, partition) work items fill every SM the same number of times
with ThreadPoolExecutor(16) as ex:
    fa = list(ex.map(lambda i: kfsynth.synth_fasta(SEED, i, 5_000_000), range(G)))
for k in (7, 8, 9, 10, 12):
    n = G if k <= 10 else 24
    arena = engine.DeviceArena(fa[:n])
    V = engine.vocab_size(k)
    if k == 12:
        counts = torch.empty((n, V), dtype=torch.int64, device="cuda")
        ms = []
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); engine.count_device(arena, k=k, counts=counts); e1.record(); torch.cuda.synchronize()
            ms.append((e0.elapsed_time(e1), engine.last_count_kernel_ms()))
        step, kern = min(m[0] for m in ms[1:]), min(m[1] for m in ms[1:])
    else:
        step, kern, counts = timed(arena, k, n, V, reps=4)
    dt, ref = cpu(fa[:4], k)
    ok = bool(np.array_equal(ref, counts[:4].cpu().numpy().astype(np.uint64)))
    alg = arena.file_bytes + n * V * 12
    emit(config="configs[4] large-k sweep, %d x 5 Mbp genomes, whole-genome mode" % n, k=k, gbases_per_s=n * 5e6 / step / 1e6, kernel_ms=kern,
         step_ms=step, roofline_frac=(alg / kern / 1e6) / PEAK, histogram="shared memory" if k <= 7 else ("partitioned shared memory, %d partition(s) per file, text parsed once (u32 row per file)" % {8: 1, 9: 3, 10: 11}[k] if k <= 10 else "global atomics (u32 row per file)"),
         cpu_gbases_per_s=4 * 5e6 / dt / 1e9, cpu_threads=threads, cpu_sample="4 genomes, oracle/kf_oracle.c", parity_ok=ok)
    del arena, counts

# ---- config 5: chunked-genome mode (10-kbp windows), k = 7 ----
import tempfile
t_plan = t_lib = t_write = 0.0
nwin = 0
kern = 0.0
tmpd = tempfile.mkdtemp()
for i in range(8):
    data = fa[i]
    t0 = time.perf_counter()
    seq, offs, lens, labels = chunks.plan_genome("g%d" % i, data)
    t1 = time.perf_counter()
    counts, _, _ = engine.count_windows(seq, offs, lens, k=7)
    t2 = time.perf_counter()
    engine.write_kf_rows(os.path.join(tmpd, "g%d.kf" % i), labels, counts.astype(np.float64), int_modes=(counts > 0).all(axis=1).astype(np.uint8))
    t3 = time.perf_counter()
    if i:
        t_plan += t1 - t0; t_lib += t2 - t1; t_write += t3 - t2; nwin += len(labels); kern += engine.last_count_kernel_ms()
# parity of the last genome's windows against the oracle's chunk rows
import kf_oracle as o
ref_rows = o.chunk_rows("g7", fa[7].tobytes(), 7)
ok = len(ref_rows) == len(labels) and all(np.array_equal(counts[j], ref_rows[j][1]) for j in range(0, len(labels), 37))
emit(config="configs[4] chunked-genome mode, 10-kbp windows, 7 x 5 Mbp genomes", k=7, windows=nwin, host_plan_ms_per_genome=t_plan / 7 * 1e3,
     library_call_ms_per_genome=t_lib / 7 * 1e3, counting_kernels_ms_per_genome=kern / 7, write_kf_ms_per_genome=t_write / 7 * 1e3,
     gbases_per_s_end_to_end=7 * 5e6 / (t_plan + t_lib + t_write) / 1e9,
     reference_note="the reference runs one jellyfish count+dump pair per window (~64 ms each in its toy log)", parity_ok=bool(ok))
