#!/usr/bin/env python
"""end to end, files on disk -> .kf files written, for N synthetic 5 Mbp genomes (page-cached input),
through kf_files_to_kf (the call get_frequencies makes).  One JSON line."""
import json, os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
from kf2vecfsw_b200 import engine
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kfsynth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
NB = 5_000_000
engine.init(0)
root = tempfile.mkdtemp(prefix="kf_files_")
ind, outd = os.path.join(root, "in"), os.path.join(root, "out")
os.makedirs(ind); os.makedirs(outd)
def gen(i):
    p = os.path.join(ind, "g%05d.fna" % i)
    kfsynth.synth_fasta(20261018, i, NB).tofile(p)
    return p
threads = len(os.sched_getaffinity(0))
with ThreadPoolExecutor(threads) as ex:
    paths = list(ex.map(gen, range(N)))
names = ["g%05d" % i for i in range(N)]
outs = [os.path.join(outd, s + ".kf") for s in names]
nbytes = sum(os.path.getsize(p) for p in paths)
res = []
for it in range(3):
    t0 = time.perf_counter()
    status, totals, secs = engine.files_to_kf(paths, outs, names, k=7, threads=threads)
    dt = time.perf_counter() - t0
    assert (status == 0).all()
    res.append((dt, secs.tolist()))
dt, secs = min(res)
out_bytes = sum(os.path.getsize(p) for p in outs)
print(json.dumps({"config": "files on disk (page cache) -> .kf files, %d x 5 Mbp genomes, k=7, kf_files_to_kf" % N, "gbases_per_s": N * NB / dt / 1e9,
                  "seconds": dt, "input_GB_per_s": nbytes / dt / 1e9, "input_bytes": nbytes, "kf_bytes_written": out_bytes, "host_threads": threads,
                  "stage_seconds": {"wait_reads": secs[0], "gpu_stage": secs[1], "wait_writes": secs[2], "total": secs[3]},
                  "first_call_seconds": res[0][0]}))
shutil.rmtree(root)
