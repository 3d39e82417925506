#!/usr/bin/env python
"""bench.py -- k=7 k-mer frequency throughput on synthetic bacterial genomes (BASELINE.json config 2 / 3).

A step = one pass of the hot path (count kernel + fold/normalise kernel; plus the NCCL all-gather of the
frequency matrix when N > 1) over one batch of synthetic genomes:
  * `value`   : inputs resident in HBM, CUDA events on the launching stream, max over ranks
  * `e2e`     : the same batch through the C-ABI host call kf_count_buffers (pinned host buffers in,
                H2D + kernels + D2H of counts/frequencies inside the timed region)
  * `roofline`: algorithmic bytes of the counting kernel / its CUDA-event duration vs the measured HBM peak
  * `cpu_baseline`: the oracle's multi-threaded C restatement on a bounded sample (rank 0, N = 1)
`--impl reference` times the reference's CPU path instead (Jellyfish is not in the image, so this is the
oracle's C restatement of it on all host threads).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import kfsynth  # noqa: E402  (input generator: tools/libkfsynth.so, not the product library)

METRIC = "Gbases/s k=7 k-mer frequency"
UNIT = "Gbases/s"
SEED = 20261018


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genomes", type=int, default=1000, help="genomes per GPU (weak scaling)")
    ap.add_argument("--bases", type=int, default=5_000_000)
    ap.add_argument("--k", type=int, default=7)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip the files-on-disk -> .kf end-to-end measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configs (line widths, FASTQ, large k)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --genomes per GPU; strong: --genomes-total sharded over the GPUs (BASELINE.json configs[2])")
    ap.add_argument("--genomes-total", type=int, default=10000, help="strong scaling: genomes of the whole job")
    return ap.parse_args()


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def generate_genomes(engine, ids, n_bases, threads, pinned=True):
    """Synthetic 80-column FASTA genomes (kf_synth_fasta) written into one (pinned) host buffer."""
    import numpy as np
    import torch
    sizes = [kfsynth.synth_fasta_size(SEED, g, n_bases) for g in ids]
    offs = np.zeros(len(ids) + 1, dtype=np.int64)
    offs[1:] = np.cumsum(sizes)
    host = torch.empty(int(offs[-1]), dtype=torch.uint8, pin_memory=pinned)
    hnp = host.numpy()
    views = [hnp[offs[i]:offs[i + 1]] for i in range(len(ids))]

    def work(i):
        kfsynth.synth_fasta(SEED, ids[i], n_bases, out=views[i])

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        list(ex.map(work, range(len(ids))))
    return host, views


def build_arena_chunked(engine, ids, n_bases, threads, dev, chunk=500):
    """Device arena of the synthetic genomes `ids`, generated chunk by chunk through one reused pinned buffer (10,000 genomes
    are 50 GB: they need not exist in host memory at once)."""
    import numpy as np
    import torch
    sizes = [kfsynth.synth_fasta_size(SEED, g, n_bases) for g in ids]
    arena = engine.DeviceArena.empty(sizes, [0x3E] * len(ids), device=dev)
    cap = max(sum(sizes[i:i + chunk]) for i in range(0, len(ids), chunk))
    stage = [torch.empty(cap, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for ci, c0 in enumerate(range(0, len(ids), chunk)):
        idx = list(range(c0, min(len(ids), c0 + chunk)))
        buf = stage[ci & 1].numpy()
        offs = np.zeros(len(idx) + 1, dtype=np.int64)
        offs[1:] = np.cumsum([sizes[i] for i in idx])
        views = [buf[offs[j]:offs[j + 1]] for j in range(len(idx))]
        with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
            list(ex.map(lambda j: kfsynth.synth_fasta(SEED, ids[idx[j]], n_bases, out=views[j]), range(len(idx))))
        if ci >= 2:
            torch.cuda.synchronize(dev)          # (the copies out of this staging buffer two chunks ago are done)
        for j, i in enumerate(idx):
            arena.load(i, views[j])
    torch.cuda.synchronize(dev)
    return arena


class ClockSampler:
    """SM clock and throttle reasons sampled every 2 ms over the timed region through NVML (the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; B200_PROFILING.md recipe)."""

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.device
            if visible:
                try:
                    idx = int(visible.split(",")[self.device])
                except ValueError:
                    idx = self.device
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                    "sw_power_cap": 0x4, "hw_power_brake_slowdown": 0x80}

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for name, bit in bits.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception as e:  # pragma: no cover
                        self.err = repr(e)
                        return
                    time.sleep(0.002)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception as e:
            self.err = repr(e)

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=2)
        sm = sorted(self.samples)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(sm)}
        if self.err:
            out["error"] = self.err
        return out


def cpu_reference_run(engine, n_bases, k, steps, warmup, budget_s_per_step=2.0):
    """The reference's CPU path (restated in oracle/kf_oracle.c) on all host threads, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    threads = host_threads()
    probe = [kfsynth.synth_fasta(SEED, 0, n_bases)]
    t0 = time.perf_counter()
    c_oracle.count_buffers_mt(probe, k, 1)
    t1 = max(time.perf_counter() - t0, 1e-4)
    per_thread = max(1, int(budget_s_per_step / t1))
    S = int(min(threads * per_thread, 1024))
    _, views = generate_genomes(engine, list(range(S)), n_bases, threads, pinned=False)
    for _ in range(warmup):
        c_oracle.count_buffers_mt(views, k, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        c_oracle.count_buffers_mt(views, k, threads)
    dt = time.perf_counter() - t0
    gbases = S * n_bases * steps / dt / 1e9
    return gbases, dt / steps * 1e3, threads, S


def workload_name(genomes, bases, k, world):
    return ("%d synthetic bacterial-size genomes per GPU x %d bases (80-col FASTA, 1-50 contigs, 10 N-runs), k=%d; "
            "BASELINE.json configs[%d]" % (genomes, bases, k, 1 if world == 1 else 2))


def common_config(genomes, bases, k, world, scaling="weak", total=None):
    """The `config` object: identical for both arms (`--impl ours` and `--impl reference`) of one launch."""
    c = {"workload": workload_name(genomes, bases, k, world), "genomes_per_gpu": genomes, "bases_per_genome": bases, "k": k,
         "l2_policy": "inputs (%.2f GB per GPU) are larger than the 126 MB L2; no flush needed" % (genomes * bases * 1.0126 / 1e9),
         "parallelism": "genome-sharded, one process per GPU, no collective on the counting path"
                        + ("; all-gather of the [N,8192] fp32 matrix of every step inside the timed region" if world > 1 else "")}
    if scaling == "strong":
        c["workload"] = ("%d synthetic bacterial-size genomes x %d bases (80-col FASTA), k=%d, sharded over %d GPU(s), [%d, 8192] fp32 matrix "
                         "gathered in file order; BASELINE.json configs[2]" % (total, bases, k, world, total))
        c["genomes_total"] = total
    return c


def peak_hbm():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def other_configs(engine, dev, threads, steps=5):
    """BASELINE.json configs[3] (FASTQ), configs[4] (large k) and the k = 7 line-width variants, each with its own roofline
    (algorithmic bytes / CUDA-event time of the counting kernels) and CPU baseline (C oracle, all host threads, bounded
    sample).  Inputs resident in HBM; rank 0 of a 1-GPU run only."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    peak, _ = peak_hbm()
    out = {}

    def timed_dense(arena, k, n, want_freq=True):
        V = engine.vocab_size(k)
        counts = torch.empty((n, V), dtype=torch.int64, device=dev)
        freq = torch.empty((n, V), dtype=torch.float64, device=dev) if want_freq else None
        ms, km = [], []
        for _ in range(steps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); engine.count_device(arena, k=k, counts=counts, freq=freq); e1.record(); torch.cuda.synchronize(dev)
            ms.append(e0.elapsed_time(e1)); km.append(engine.last_count_kernel_ms())
        return sum(ms[1:]) / steps, sum(km[1:]) / steps, counts, V

    def cpu(bufs, k):
        t0 = time.perf_counter()
        ref, _, _ = c_oracle.count_buffers_mt(bufs, k, threads, want_freq=True)
        return time.perf_counter() - t0, ref

    def entry(name, workload, bases, step_ms, kern_ms, alg_bytes, kernel, cpu_bases, cpu_s, cpu_sample, ok, **kw):
        out[name] = dict({"workload": workload, "value": bases / step_ms / 1e6, "unit": UNIT, "ms_per_step": step_ms,
                          "roofline": {"bound": "hbm", "achieved": alg_bytes / kern_ms / 1e6, "peak": peak, "unit": "GB/s",
                                       "frac": alg_bytes / kern_ms / 1e6 / peak, "kernel": kernel, "kernel_ms": kern_ms,
                                       "algorithmic_bytes_per_launch": int(alg_bytes)},
                          "roofline_step": {"frac": alg_bytes / step_ms / 1e6 / peak, "window": "memsets + probe + counting + fold/normalise"},
                          "cpu_baseline": {"value": cpu_bases / cpu_s / 1e9, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_sample},
                          "parity_ok": bool(ok)}, **kw)

    # ---- k = 7, other line structures: unwrapped (one line per contig), 100 and 60 columns ----
    for name, lw, G7, kern in (("k7_unwrapped", 10 ** 9, 1000, "count_fasta_lines_kernel, virtual lines of 80 bytes (vl_process_piece)"),
                               ("k7_100col", 100, 500, "count_fasta_lines_kernel, ln_process_piece<100>"),
                               ("k7_60col", 60, 500, "count_fasta_lines_kernel, ln_process_piece<60>")):
        with ThreadPoolExecutor(threads) as ex:
            fa = list(ex.map(lambda i: kfsynth.synth_fasta(SEED, i, 5_000_000, line_width=lw), range(G7)))
        arena = engine.DeviceArena(fa, device=dev)
        step_ms, kern_ms, counts, V = timed_dense(arena, 7, G7)
        dt, ref = cpu(fa[:threads * 2], 7)
        ok = np.array_equal(ref[:4], counts[:4].cpu().numpy().astype(np.uint64))
        entry(name, "%d synthetic genomes x 5 Mbp, k=7, %s" % (G7, "unwrapped FASTA (one line per contig)" if lw > 1000 else "%d-column FASTA" % lw),
              G7 * 5e6, step_ms, kern_ms, arena.file_bytes + G7 * V * 12, kern, threads * 2 * 5e6, dt, "%d of the genomes, oracle/kf_oracle.c" % (threads * 2), ok)
        del arena, counts, fa
    # ---- plan-cache miss: a stream of DISTINCT batches (every call a new layout: tile plan built on the host, tables
    # uploaded stream-ordered) against the same batch repeated (plan cache hit, what the timed loop above measures) ----
    with ThreadPoolExecutor(threads) as ex:
        fa = list(ex.map(lambda i: kfsynth.synth_fasta(SEED, i, 5_000_000), range(1000)))
    arenas = [engine.DeviceArena(fa[:500], device=dev), engine.DeviceArena(fa[500:], device=dev)]
    V7 = engine.vocab_size(7)
    cnt = torch.empty((500, V7), dtype=torch.int64, device=dev)
    frq = torch.empty((500, V7), dtype=torch.float64, device=dev)
    def run_seq(seq):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for a in seq:
            engine.count_device(arenas[a], k=7, counts=cnt, freq=frq)
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / len(seq) * 1e3
    run_seq([0, 1, 0, 1])
    hit_ms = min(run_seq([0] * 10), run_seq([1] * 10))
    miss_ms = run_seq([0, 1] * 5)
    out["plan_cache"] = {"workload": "500 genomes x 5 Mbp per call, k=7, wall clock per call incl. host work (calls enqueue back to back)",
                         "hit_ms_per_call": hit_ms, "miss_ms_per_call": miss_ms, "miss_cost_ms": miss_ms - hit_ms,
                         "note": "a miss builds the tile plan on the host (~1,100 tiles) and uploads six small tables through pinned staging on the launching stream; no device synchronisation"}
    del arenas, cnt, frq, fa
    # ---- configs[3]: FASTQ query reads, 150 bp, 30x of 5 Mbp, N-containing ----
    n_samples, n_reads = 8, 1_000_000
    with ThreadPoolExecutor(threads) as ex:
        fq = list(ex.map(lambda i: kfsynth.synth_fastq(SEED, i, 5_000_000, n_reads, 150), range(n_samples)))
    arena = engine.DeviceArena(fq, device=dev)
    step_ms, kern_ms, counts, V = timed_dense(arena, 7, n_samples)
    dt, ref = cpu(fq[:2], 7)
    ok = np.array_equal(ref, counts[:2].cpu().numpy().astype(np.uint64)) and bool((engine.last_file_status(arena) == 0).all())
    entry("fastq_k7", "BASELINE.json configs[3]: %d FASTQ samples x %d reads x 150 bp (30x of 5 Mbp, N-containing), k=7" % (n_samples, n_reads),
          n_samples * n_reads * 150, step_ms, kern_ms, arena.file_bytes + n_samples * V * 12, "count_fastq_pairs_kernel (8-mer pair histogram; every lane chases its own records)",
          2 * n_reads * 150, dt, "2 of the samples, oracle/kf_oracle.c", ok, bytes_per_base=arena.file_bytes / (n_samples * n_reads * 150))
    del arena, counts, fq
    # ---- configs[4]: large-k sweep on 5 Mbp genomes ----
    GL = 296
    with ThreadPoolExecutor(threads) as ex:
        fa = list(ex.map(lambda i: kfsynth.synth_fasta(SEED, i, 5_000_000), range(GL)))
    arena = engine.DeviceArena(fa, device=dev)
    for k in (8, 9, 10):
        step_ms, kern_ms, counts, V = timed_dense(arena, k, GL, want_freq=False)
        dt, ref = cpu(fa[:4], k)
        ok = np.array_equal(ref, counts[:4].cpu().numpy().astype(np.uint64))
        entry("large_k%d" % k, "BASELINE.json configs[4]: %d synthetic genomes x 5 Mbp, k=%d, dense canonical rows" % (GL, k), GL * 5e6, step_ms,
              kern_ms, arena.file_bytes + GL * V * 12, "count_fasta_part_kernel<%d> (partitioned shared-memory histogram) + tiled fold" % k,
              4 * 5e6, dt, "4 of the genomes, oracle/kf_oracle.c", ok)
        del counts
    # k = 11, 12: sparse (observed canonical k-mers, sorted) -- the sort-and-run-length path on the same 296 genomes; parity
    # of two genomes' entries against the C oracle (fetched from a small arena of their own)
    arena_s = engine.DeviceArena(fa[:2], device=dev)
    for k in (11, 12):
        ms = []
        for _ in range(3):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            _, _, row_off, _, _ = engine.sparse_count_device(arena, k, fetch=False)
            torch.cuda.synchronize(dev)
            ms.append((time.perf_counter() - t0) * 1e3)
        entries = int(row_off[-1])
        engine.sparse_release()
        codes, cnts, row_off2, totals, _ = engine.sparse_count_device(arena_s, k)
        t0 = time.perf_counter()
        rc, rn, rt = c_oracle.count_sparse(fa[0].tobytes(), k)
        dt = time.perf_counter() - t0
        a, e = int(row_off2[0]), int(row_off2[1])
        ok = rt == int(totals[0]) and np.array_equal(codes[a:e], rc) and np.array_equal(cnts[a:e].astype(np.uint64), rn) and \
            int(row_off[1] - row_off[0]) == e - a
        engine.sparse_release()
        step_ms = min(ms[1:])
        name = "large_k%d_sparse" % k
        entry(name, "BASELINE.json configs[4]: %d synthetic genomes x 5 Mbp, k=%d, sparse (code, count) output sorted by code" % (GL, k),
              GL * 5e6, step_ms, step_ms, arena.file_bytes + entries * 12, "sparse_tile / sparse_wc_scatter / sparse16_distinct / sparse16_emit "
              "(whole call, host-timed: it synchronises once per sub-batch to size the output)", 5e6, dt,
              "1 genome, 1 thread (sort + run lengths), oracle/kf_oracle.c", ok, entries=entries)
        out[name]["cpu_baseline"]["cores"] = 1
    del arena, arena_s, fa
    return out


# SMs left to the overlapped all-gather when N > 1 (measured on 8 B200: 8 CTAs move the 262 MB gather in about the time
# of one counting step, 4 are too few; 2 GPUs exchange a quarter of that)
NCCL_CTAS = int(os.environ.get("KF_BENCH_NCCL_CTAS", "0"))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from kf2vecfsw_b200 import engine

    if args.impl == "reference":
        if rank != 0:
            return 0
        gb, ms, threads, S = cpu_reference_run(engine, args.bases, args.k, args.steps, args.warmup)
        sample = "%d synthetic genomes x %d bases per step (same generator/config as the GPU arm)" % (S, args.bases)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": gb, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 counts / f64 frequencies", "data": "synthetic",
            "config": common_config(args.genomes, args.bases, args.k, args.gpus),
            "reference_arm": "CPU restatement of jellyfish count -C + dump -c + vocabulary merge + normalise (oracle/kf_oracle.c; the "
                             "Jellyfish binary is not in the image), all host threads, each step a bounded sample of the workload: " + sample,
            "cpu_baseline": {"value": gb, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": gb, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from kf2vecfsw_b200 import dist as kfdist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    global NCCL_CTAS
    if NCCL_CTAS <= 0:
        NCCL_CTAS = 8
    # How the [N, V] matrix is assembled when N > 1.  "peer" (default): every rank pushes its block into every peer's matrix
    # with device-to-device copies over NVLink (copy engines, no SM: kf2vecfsw_b200.dist.PeerGather), the gather of step i
    # beside the counting of step i + 1, which keeps all SMs.  "nccl": NCCL's all-gather (round 1), overlapped from 4 GPUs
    # on with NCCL_MAX_CTAS CTAs and the counting kernels sized for the remaining SMs.
    gather_impl = os.environ.get("KF_BENCH_GATHER", "peer") if world > 1 else None
    overlap = gather_impl == "nccl" and (world >= 4 or "KF_BENCH_NCCL_CTAS" in os.environ)
    if overlap:
        os.environ.setdefault("NCCL_MAX_CTAS", str(NCCL_CTAS))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    engine.init(local_rank)
    numa_node = engine.bind_host_to_gpu(local_rank) if world > 1 else None   # before any pinned host allocation
    sms_used = engine.set_sm_limit(0)
    if overlap:
        sms_used = engine.set_sm_limit(sms_used - NCCL_CTAS)
    k, NB = args.k, args.bases
    strong = args.scaling == "strong"
    if strong:
        # BASELINE.json configs[2]: --genomes-total genomes sharded over the GPUs in contiguous blocks (rank order = file
        # order, so the gathered matrix is in the original row order)
        tot = args.genomes_total
        rows = [tot // world + (1 if r < tot % world else 0) for r in range(world)]
    else:
        rows = [args.genomes] * world
    G = rows[rank]
    first = sum(rows[:rank])
    V = engine.vocab_size(k)
    threads = max(1, host_threads() // max(1, world))

    ids = list(range(first, first + G))
    host = views = None
    if strong:
        arena = build_arena_chunked(engine, ids, NB, threads, dev)
        args.no_e2e = True            # (the end-to-end arm is measured on the weak-scaling configuration)
    else:
        host, views = generate_genomes(engine, ids, NB, threads)
        arena = engine.DeviceArena(views, device=dev)
    file_bytes = arena.file_bytes
    counts = torch.empty((G, V), dtype=torch.int64, device=dev)
    freq = torch.empty((G, V), dtype=torch.float64, device=dev)
    feat = torch.empty((G, V), dtype=torch.float32, device=dev)
    totals = torch.empty(G, dtype=torch.int64, device=dev)
    pg = kfdist.PeerGather(rows, V, torch.float32, dev) if gather_impl == "peer" else None
    og = kfdist.OverlappedGather(G, V, torch.float32, dev) if overlap else None
    gathered = torch.empty((world * max(rows), V), dtype=torch.float32, device=dev) if (gather_impl == "nccl" and not overlap) else None
    gpad = torch.zeros((max(rows), V), dtype=torch.float32, device=dev) if gathered is not None and G < max(rows) else None
    kernel_ms = []
    nsub = [0]

    def step(record=False):
        # everything is enqueued on torch's current stream: no host synchronisation inside a step
        f = pg.slot() if pg else (og.slot() if og else feat)
        engine.count_device(arena, k=k, counts=counts, freq=freq, feat=f, totals=totals)
        if pg:
            pg.submit()
            if nsub[0] > 0:            # the previous step's matrix: complete by now (its pushes ran beside this counting)
                pg.wait(step=nsub[0] - 1)
                pg.release(step=nsub[0] - 1)
            nsub[0] += 1
        elif og:
            og.submit()
        elif gathered is not None:
            if gpad is not None:
                gpad[:G].copy_(f)
                dist.all_gather_into_tensor(gathered, gpad)
            else:
                dist.all_gather_into_tensor(gathered, f)
        if record:
            kernel_ms.append(engine.last_count_kernel_ms())   # (waits for the library's events: only outside the timed region)

    def drain():
        if pg and nsub[0] > 0:
            pg.wait(step=nsub[0] - 1)
            pg.release(step=nsub[0] - 1)
        if og:
            og.drain()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    drain()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    drain()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # duration of the counting kernels alone, OVER THE TIMED REGION: the library records a pair of CUDA events around them
    # on the launching stream at every call (a ring of 64 pairs), read back only now
    hist = engine.count_kernel_ms_history(min(args.steps, 64))
    kernel_ms.extend(float(x) for x in hist)
    if not kernel_ms:
        for _ in range(min(5, args.steps)):
            step(record=True)
    barrier()
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    kms = torch.tensor([sum(kernel_ms) / len(kernel_ms)], dtype=torch.float64, device=dev)
    launches = engine.last_launch_count() * args.steps
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total.item()) / args.steps
    value = sum(rows) * NB / (ms_step * 1e-3) / 1e9
    gather_ok = None
    if pg:
        # the gathered matrix of the last step against this rank's own rows and a checksum of every rank's block
        full = pg.full[(nsub[0] - 1) % pg.NBUF]
        mine = bool(torch.equal(full[first:first + G], (freq * 1e4).to(torch.float32)))
        sums = torch.zeros(world, dtype=torch.float64, device=dev)
        sums[rank] = full[first:first + G].double().sum()
        dist.all_reduce(sums)
        got = torch.stack([full[sum(rows[:r]):sum(rows[:r + 1])].double().sum() for r in range(world)])
        gather_ok = mine and bool(torch.equal(got, sums))

    # parity spot check against the oracle (not timed): first and last genome of this rank
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    pick = [0, G - 1] if G > 1 else [0]
    pick_bufs = [views[i] for i in pick] if views is not None else [kfsynth.synth_fasta(SEED, ids[i], NB) for i in pick]
    ref, _, _ = c_oracle.count_buffers_mt(pick_bufs, k, 2, want_freq=False)
    got = counts[pick].cpu().numpy().astype(np.uint64)
    parity_ok = bool(np.array_equal(ref, got))

    # end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        out_counts = torch.empty((G, V), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
        out_freq = torch.empty((G, V), dtype=torch.float64, pin_memory=True).numpy()
        for _ in range(max(1, args.warmup)):
            engine.count_buffers(views, k=k, out_counts=out_counts, out_freq=out_freq)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            engine.count_buffers(views, k=k, out_counts=out_counts, out_freq=out_freq)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_ok = bool(np.array_equal(out_counts[pick], ref))
        e2e = {"value": sum(rows) * NB * args.steps / float(dt.item()) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(file_bytes), "d2h_bytes_per_step": int(G * V * 16 + G * 8),
               "ms_per_step": float(dt.item()) / args.steps * 1e3, "parity_ok": e2e_ok,
               "api": "kf_count_buffers (C ABI, pinned host buffers in, counts+frequencies out)",
               "host_numa_node_rank0": numa_node}
        hp = os.path.join(ROOT, "profiles", "h2d_ceiling.json")
        if os.path.exists(hp):
            hc = json.load(open(hp)).get(str(world))
            if hc:   # what the box's host-to-device copies alone reach with this many ranks copying at once
                e2e["bound_gbs"] = hc["h2d_concurrent_total_gbs"]
                e2e["bound"] = "concurrent pinned host-to-device copies, no kernels: " + hc["source"]
                e2e["h2d_gbs_achieved"] = e2e["h2d_bytes_per_step"] * world / (e2e["ms_per_step"] * 1e-3) / 1e9

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gb, ms, cthreads, S = cpu_reference_run(engine, NB, k, steps=3, warmup=1, budget_s_per_step=1.5)
        cpu_baseline = {"value": gb, "unit": UNIT, "cores": cthreads, "kind": "port",
                        "sample": "%d of the same synthetic genomes per pass, 3 timed passes, oracle/kf_oracle.c "
                                  "(rolling canonical counter + normalise), one genome per thread" % S}

    # files on disk -> .kf files written, through the call get_frequencies makes (kf_files_to_kf: reads,
    # GPU stage and writes pipelined), on a bounded number of the same genomes written to a scratch directory first
    e2e_files = None
    if rank == 0 and world == 1 and not args.no_e2e and not args.no_files and views is not None:
        import shutil
        import tempfile
        NF = min(G, 400)
        root = tempfile.mkdtemp(prefix="kf_bench_files_")
        try:
            ind, outd = os.path.join(root, "in"), os.path.join(root, "out")
            os.makedirs(ind)
            os.makedirs(outd)
            names = ["g%05d" % ids[i] for i in range(NF)]
            paths = [os.path.join(ind, s + ".fna") for s in names]
            for i in range(NF):
                views[i].tofile(paths[i])
            outs = [os.path.join(outd, s + ".kf") for s in names]
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                st, _, secs = engine.files_to_kf(paths, outs, names, k=k, threads=threads)
                dtf = time.perf_counter() - t0
                if best is None or dtf < best[0]:
                    best = (dtf, secs.tolist())
            row0 = open(outs[0]).read().rstrip("\n").split(",")
            files_ok = bool((st == 0).all()) and row0[0] == names[0] and len(row0) == V + 1 and \
                bool(np.array_equal(np.array(row0[1:], dtype=np.float64), freq[0].cpu().numpy()))
            e2e_files = {"value": NF * NB / best[0] / 1e9, "unit": UNIT, "files": NF, "seconds": best[0],
                         "input_bytes": int(sum(v.size for v in views[:NF])), "kf_bytes_written": int(sum(os.path.getsize(o) for o in outs)),
                         "host_threads": threads, "stage_seconds": dict(zip(("wait_reads", "gpu_stage", "wait_writes", "total"), best[1])),
                         "parity_ok": files_ok, "api": "kf_files_to_kf (what get_frequencies calls): page-cached .fna files in, .kf text files out"}
        finally:
            shutil.rmtree(root, ignore_errors=True)

    configs = None
    if rank == 0 and world == 1 and not args.no_configs and not strong:
        del arena, counts, freq, feat, host
        views = None
        torch.cuda.empty_cache()
        configs = other_configs(engine, dev, threads)

    if rank == 0:
        peak, peak_src = peak_hbm()
        alg_bytes = file_bytes + G * V * 4 + G * V * 8          # BASELINE.md section 3, per rank
        # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from one `ncu --set full` capture of this
        # command (tools/ncu_summary.py writes the file); only quoted when it was taken on the same workload
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            t = json.load(open(tpath))
            if t.get("genomes") == G and t.get("bases") == NB and t.get("k") == k:
                traffic, traffic_src = t["dram_bytes_per_launch"], t["source"]
        achieved = alg_bytes / (float(kms.item()) * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u32 shared-memory counts -> u64 counts, f64 frequencies", "data": "synthetic",
            "config": common_config(G if not strong else rows[0], NB, k, world, args.scaling, sum(rows)),
            "file_bytes_per_gpu": int(file_bytes),
            "gather": None if world == 1 else (
                "peer pushes over NVLink copy engines (cuMemcpyDtoDAsync into every peer's [N,V] matrix + flag words waited on with "
                "cuStreamWaitValue32: no SM, no NCCL kernel), the gather of step i beside the counting of step i+1 on all %d SMs" % sms_used
                if pg else ("NCCL all-gather of step i overlapping the counting of step i+1 (two buffer pairs; NCCL_MAX_CTAS=%d, counting "
                            "kernels sized for %d of the SMs)" % (NCCL_CTAS, sms_used) if overlap else "NCCL all-gather after the counting")),
            "gather_ok": gather_ok,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "count_fasta_lines_kernel (one launch for all line widths; ln_process_piece<80> on this input)",
                         "window": "width probe + counting kernels (CUDA events on the launching stream inside the library; the launches for "
                                   "other widths and the generic kernel exit at once on this input); the fold/normalise kernel is NOT in this "
                                   "window -- see roofline_step",
                         "kernel_ms": float(kms.item()), "algorithmic_bytes_per_launch": int(alg_bytes), "peak_source": peak_src},
            "roofline_step": {"bound": "hbm", "achieved": alg_bytes / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": alg_bytes / (ms_step * 1e-3) / 1e9 / peak, "ms": ms_step,
                              "window": "the whole step as timed for `value`: probe + counting + fold/normalise (BASELINE.md section 3)"
                                        + (" + all-gather" if world > 1 else "")},
            "configs": configs,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "e2e_files": e2e_files,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity_ok": parity_ok,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
