"""Multi-GPU plumbing for the k-mer frequency step: genomes are independent units (the loop body of
kf2vec/main.py:301 touches only file i), so the counting path shards with NO collective.  The only exchange is
assembling the per-rank [n_r, V] float32 feature blocks (fp32(freq*1e4), train_classifier_model.py:149) into the
[N, V] backbone matrix -- one all-gather over NCCL/NVLink (gloo on CPU for tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_by_size(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-processing-time partition of file indices by byte size; deterministic.
    Returns world lists of indices; within a rank indices stay in increasing order."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    loads = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (loads[j], j))
        parts[r].append(i)
        loads[r] += int(sizes[i])
    for p in parts:
        p.sort()
    return parts


def all_gather_rows(local, parts: List[List[int]], group=None):
    """local: [len(parts[rank]), V] tensor on this rank.  Returns the [N, V] matrix in ORIGINAL file order on every
    rank.  Uneven shards are padded to the largest shard for all_gather_into_tensor and unpadded afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert local.shape[0] == len(parts[rank])
    V = local.shape[1]
    m = max(len(p) for p in parts)
    if local.shape[0] < m:
        pad = torch.zeros((m - local.shape[0], V), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    out = torch.empty((world * m, V), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    N = sum(len(p) for p in parts)
    full = torch.empty((N, V), dtype=local.dtype, device=local.device)
    for r, p in enumerate(parts):
        if p:
            idx = torch.tensor(p, dtype=torch.long, device=local.device)
            full[idx] = out[r * m: r * m + len(p)]
    return full


class OverlappedGather:
    """The all-gather of batch i runs while batch i + 1 is being counted: two (local, gathered) buffer pairs used in
    turn; ``slot()`` hands out the next local block to fill (after waiting for the gather that last read it),
    ``submit()`` starts its all-gather without blocking, ``drain()`` orders every outstanding gather before whatever the
    caller enqueues next.  With NCCL the waits are stream dependencies (no host synchronisation); with gloo (CPU tests)
    they block."""

    def __init__(self, rows: int, cols: int, dtype, device, group=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.local = [torch.empty((rows, cols), dtype=dtype, device=device) for _ in range(2)]
        self.full = [torch.empty((self.world * rows, cols), dtype=dtype, device=device) for _ in range(2)]
        self.pending = [None, None]
        self.i = 0

    def slot(self):
        b = self.i & 1
        if self.pending[b] is not None:
            self.pending[b].wait()
            self.pending[b] = None
        return self.local[b]

    def submit(self):
        """Starts the all-gather of the block last handed out by slot(); returns the [world * rows, cols] tensor it
        fills (valid after drain(), or after the slot() call two batches later)."""
        import torch.distributed as dist
        b = self.i & 1
        self.pending[b] = dist.all_gather_into_tensor(self.full[b], self.local[b], group=self.group, async_op=True)
        self.i += 1
        return self.full[b]

    def drain(self):
        for b in (0, 1):
            if self.pending[b] is not None:
                self.pending[b].wait()
                self.pending[b] = None
