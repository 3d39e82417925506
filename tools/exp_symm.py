#!/usr/bin/env python
"""Peer copy bandwidth into a torch symmetric-memory buffer (cuMem / fabric handles) vs CUDA-IPC: torchrun --nproc-per-node 2 tools/exp_symm.py"""
import json, os, sys
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rows, V = 1000, 8192
res = {}
try:
    t = symm_mem.empty((world * rows, V), dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    peer = hdl.get_buffer((rank + 1) % world, (world * rows, V), torch.float32)
    src = torch.ones((rows, V), device=dev) * (rank + 1)
    for name, fn in (("symm_torch_copy", lambda: peer[rank * rows:(rank + 1) * rows].copy_(src, non_blocking=True)),):
        fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        res[name + "_gbs"] = rows * V * 4 * 10 / e0.elapsed_time(e1) / 1e6
    dist.barrier(); torch.cuda.synchronize()
    other = (rank - 1) % world
    res["ok"] = bool((t[other * rows:(other + 1) * rows] == float(other + 1)).all())
    res["signal_pad"] = str(hdl.get_signal_pad(rank).shape) + str(hdl.get_signal_pad(rank).dtype)
except Exception as e:
    import traceback
    res["error"] = repr(e); traceback.print_exc()
if rank == 0:
    print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()
