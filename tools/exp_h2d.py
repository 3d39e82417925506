#!/usr/bin/env python
"""Host-to-device copy ceiling of the box: every rank copies a pinned 2 GiB buffer to its GPU, first one rank at a time, then
all ranks at once (no kernels).  torchrun --nproc-per-node N tools/exp_h2d.py  -> one JSON line (rank 0)."""
import json, os, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
nbytes = 2 << 30
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
h.fill_(rank)
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
def run(reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return nbytes * reps / e0.elapsed_time(e1) / 1e6   # GB/s
run(1)
solo = torch.zeros(world, dtype=torch.float64, device=dev)
for r in range(world):
    dist.barrier(); torch.cuda.synchronize()
    if r == rank:
        solo[r] = run()
dist.barrier(); torch.cuda.synchronize()
both = torch.zeros(world, dtype=torch.float64, device=dev)
both[rank] = run()
dist.all_reduce(solo); dist.all_reduce(both)
if rank == 0:
    print(json.dumps({"world": world, "h2d_solo_gbs": [round(x, 1) for x in solo.tolist()], "h2d_concurrent_gbs": [round(x, 1) for x in both.tolist()],
                      "h2d_concurrent_total_gbs": round(float(both.sum()), 1), "cpus": len(os.sched_getaffinity(0)),
                      "bases_per_s_ceiling_at_1.0126_bytes_per_base": round(float(both.sum()) / 1.0126, 1)}))
dist.barrier(); dist.destroy_process_group()
