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
            "config": {"workload": workload_name(args.genomes, args.bases, args.k, args.gpus),
                       "genomes_per_gpu": args.genomes, "bases_per_genome": args.bases, "k": args.k,
                       "reference_arm": "CPU restatement of jellyfish count -C + dump -c + vocabulary merge + normalise "
                                        "(oracle/kf_oracle.c; the Jellyfish binary is not in the image), all host threads, "
                                        "each step a bounded sample of the workload: " + sample},
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
    # measured: at 2 GPUs the 65 MB gather costs less (0.07 ms) than the SMs the overlap needs; from 4 GPUs on it pays
    overlap = world >= 4 or (world > 1 and "KF_BENCH_NCCL_CTAS" in os.environ)
    if overlap:
        # the all-gather of batch i runs beside the counting of batch i+1: NCCL gets at most NCCL_CTAS CTAs, and the
        # counting kernels (one persistent CTA per SM, ~205 KB of shared memory each) are sized for the other SMs
        os.environ.setdefault("NCCL_MAX_CTAS", str(NCCL_CTAS))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    engine.init(local_rank)
    numa_node = engine.bind_host_to_gpu(local_rank) if world > 1 else None   # before any pinned host allocation
    sms_used = engine.set_sm_limit(0)
    if overlap:
        sms_used = engine.set_sm_limit(sms_used - NCCL_CTAS)
    k, G, NB = args.k, args.genomes, args.bases
    V = engine.vocab_size(k)
    threads = max(1, host_threads() // max(1, world))

    ids = list(range(rank * G, (rank + 1) * G))
    host, views = generate_genomes(engine, ids, NB, threads)
    arena = engine.DeviceArena(views, device=dev)
    file_bytes = arena.file_bytes
    counts = torch.empty((G, V), dtype=torch.int64, device=dev)
    freq = torch.empty((G, V), dtype=torch.float64, device=dev)
    feat = torch.empty((G, V), dtype=torch.float32, device=dev)
    totals = torch.empty(G, dtype=torch.int64, device=dev)
    # N > 1: the [N, V] backbone matrix is assembled on every GPU by an all-gather that runs while the next batch is
    # being counted (two buffer pairs); every gather completes inside the timed region (drain before the end event)
    og = kfdist.OverlappedGather(G, V, torch.float32, dev) if overlap else None
    gathered = torch.empty((world * G, V), dtype=torch.float32, device=dev) if (world > 1 and not overlap) else None
    kernel_ms = []

    def step(record=False):
        # everything is enqueued on torch's current stream: no host synchronisation inside a step
        f = og.slot() if og else feat
        engine.count_device(arena, k=k, counts=counts, freq=freq, feat=f, totals=totals)
        if og:
            og.submit()
        elif gathered is not None:
            dist.all_gather_into_tensor(gathered, f)
        if record:
            kernel_ms.append(engine.last_count_kernel_ms())   # (waits for the library's events: only outside the timed region)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    if og:
        og.drain()
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
    value = world * G * NB / (ms_step * 1e-3) / 1e9

    # parity spot check against the oracle (not timed): first and last genome of this rank
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    pick = [0, G - 1] if G > 1 else [0]
    ref, _, _ = c_oracle.count_buffers_mt([views[i] for i in pick], k, 2, want_freq=False)
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
        e2e = {"value": world * G * NB * args.steps / float(dt.item()) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(file_bytes), "d2h_bytes_per_step": int(G * V * 16 + G * 8),
               "ms_per_step": float(dt.item()) / args.steps * 1e3, "parity_ok": e2e_ok,
               "api": "kf_count_buffers (C ABI, pinned host buffers in, counts+frequencies out)",
               "host_numa_node_rank0": numa_node}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gb, ms, cthreads, S = cpu_reference_run(engine, NB, k, steps=3, warmup=1, budget_s_per_step=1.5)
        cpu_baseline = {"value": gb, "unit": UNIT, "cores": cthreads, "kind": "port",
                        "sample": "%d of the same synthetic genomes per pass, 3 timed passes, oracle/kf_oracle.c "
                                  "(rolling canonical counter + normalise), one genome per thread" % S}

    # files on disk -> .kf files written, through the call get_frequencies makes (kf_files_to_kf: reads,
    # GPU stage and writes pipelined), on a bounded number of the same genomes written to a scratch directory first
    e2e_files = None
    if rank == 0 and world == 1 and not args.no_e2e and not args.no_files:
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

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
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
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 shared-memory counts -> u64 counts, f64 frequencies", "data": "synthetic",
            "config": {"workload": workload_name(G, NB, k, world),
                       "genomes_per_gpu": G, "bases_per_genome": NB, "k": k, "file_bytes_per_gpu": int(file_bytes),
                       "l2_policy": "inputs (%.2f GB per GPU) are larger than the 126 MB L2; no flush needed" % (file_bytes / 1e9),
                       "parallelism": "genome-sharded, one process per GPU, no collective on the counting path"
                                      + ("; NCCL all-gather of the [N,8192] fp32 matrix of every step inside the timed region" if world > 1 else "")
                                      + (", the gather of step i overlapping the counting of step i+1 (two buffer pairs; NCCL_MAX_CTAS=%d, "
                                         "counting kernels sized for %d of the SMs)" % (NCCL_CTAS, sms_used) if overlap else "")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "count_fasta_lines_kernel<80,512> (+ width probe; the 60/70-column and generic launches "
                                   "exit at once on this input)", "kernel_ms": float(kms.item()),
                         "algorithmic_bytes_per_launch": int(alg_bytes), "peak_source": peak_src},
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
