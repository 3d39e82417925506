#!/usr/bin/env python
"""Copy-engine all-gather (dist.PeerGather) against NCCL's all-gather, no compute beside them: [rows, 8192] fp32 per rank.
torchrun --nproc-per-node N tools/exp_peer.py [rows]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from kf2vecfsw_b200 import dist as kfdist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
V = 8192
pg = kfdist.PeerGather([rows] * world, V, torch.float32, dev)
res = {}
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(20):
        pg.slot().fill_(float(rank + s))
        pg.submit()
        pg.wait()
        pg.release()
    e1.record(); torch.cuda.synchronize()
    res["peer_ms"] = e0.elapsed_time(e1) / 20
full = pg.full[(pg.t - 1) % 2]
ok = all(bool((full[r * rows:(r + 1) * rows] == float(r + 19)).all()) for r in range(world))
loc = torch.empty((rows, V), device=dev); out = torch.empty((world * rows, V), device=dev)
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(20):
        loc.fill_(1.0)
        dist.all_gather_into_tensor(out, loc)
    e1.record(); torch.cuda.synchronize()
    res["nccl_ms"] = e0.elapsed_time(e1) / 20
# one raw peer copy, timed alone
src = pg.slot()
for name, fn in (("torch_copy", lambda: pg.peer_full[(rank + 1) % world][0][rank * rows:(rank + 1) * rows].copy_(src, non_blocking=True)),
                 ("dtod", lambda: pg._copy(pg.peer_full[(rank + 1) % world][0].data_ptr() + rank * rows * V * 4, src.data_ptr(), rows * V * 4, torch.cuda.current_stream().cuda_stream))):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    res[name + "_gbs"] = rows * V * 4 * 10 / e0.elapsed_time(e1) / 1e6
if rank == 0:
    print(json.dumps({"world": world, "rows": rows, "block_mb": rows * V * 4 / 1e6, "ok": ok, **res}))
dist.barrier(); dist.destroy_process_group()
