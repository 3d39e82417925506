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


class PeerGather:
    """All-gather of the per-rank [rows_r, cols] blocks into the [N, cols] matrix on every GPU of one node WITHOUT using an
    SM.  The counting kernels are persistent CTAs that fill every SM (one CTA of ~210 KB of shared memory each), so a
    collective that runs as a kernel -- NCCL's ring -- either waits for them or needs SMs set aside for it.  Here every
    rank pushes its block straight into its slot of every peer's matrix with device-to-device copies over NVLink
    (``cuMemcpyDtoDAsync`` on a side stream: copy engines), followed by a 4-byte copy of the step number into the peer's
    flag word; a consumer orders its stream behind the arrival of all blocks with ``cuStreamWaitValue32`` on its OWN flag
    words.  No kernel, no host synchronisation in a step.  The matrices live in torch symmetric memory (set up once).

    ``slot()`` returns the view of this rank's rows inside its own matrix (the producer writes there: no local copy);
    ``submit(producer_stream)`` pushes it to the peers; ``wait(stream)`` makes ``stream`` wait for every rank's block of
    that step and returns the matrix; ``release(stream)`` tells the peers, in stream order, that the matrix may be
    overwritten (two matrices are used in turn: a push for step t+2 waits for every peer's release of step t)."""

    NBUF = 2

    def __init__(self, rows, cols: int, dtype, device, group=None):
        import torch
        import torch.distributed as dist
        from cuda.bindings import driver
        self.drv = driver
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows = [int(r) for r in rows]
        assert len(self.rows) == self.world
        self.row0 = [sum(self.rows[:r]) for r in range(self.world)]
        self.N = sum(self.rows)
        self.cols = cols
        self.device = device
        # Symmetric memory (cuMem allocations exchanged as fabric / fd handles, what NCCL itself uses): peer copies into it run
        # over NVLink at ~550 GB/s on B200; memory shared through classic CUDA IPC handles was measured at 24 GB/s here.
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.full, self.flags, self.peer_full, self.peer_flags = [], [], [[] for _ in range(self.world)], [[] for _ in range(self.world)]
        for _ in range(self.NBUF):
            t = symm_mem.empty((self.N, cols), dtype=dtype, device=device)
            h = symm_mem.rendezvous(t, group=grp)
            # flag words: [0 : world) arrival of rank q's block, [world : 2 world) rank q's release of this matrix
            f = symm_mem.empty((2 * self.world,), dtype=torch.int32, device=device)
            hf = symm_mem.rendezvous(f, group=grp)
            t.zero_()
            f.zero_()
            self.full.append(t)
            self.flags.append(f)
            for q in range(self.world):
                self.peer_full[q].append(t if q == self.rank else h.get_buffer(q, (self.N, cols), dtype))
                self.peer_flags[q].append(f if q == self.rank else hf.get_buffer(q, (2 * self.world,), torch.int32))
        self.steps = torch.arange(0, 1 << 20, dtype=torch.int32, device=device)   # source of the 4-byte flag copies (step tags)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.ready = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.t = 0            # steps submitted
        self.esize = self.full[0].element_size()

    def _check(self, res):
        err = res[0] if isinstance(res, tuple) else res
        if int(err) != 0:
            raise RuntimeError("CUDA driver call failed: %s" % err)

    def _wait_geq(self, stream_handle: int, ptr: int, value: int):
        drv = self.drv
        self._check(drv.cuStreamWaitValue32(drv.CUstream(stream_handle), drv.CUdeviceptr(ptr), value,
                                            int(drv.CUstreamWaitValue_flags.CU_STREAM_WAIT_VALUE_GEQ.value)))

    def _copy(self, dst: int, src: int, nbytes: int, stream_handle: int):
        drv = self.drv
        self._check(drv.cuMemcpyDtoDAsync(drv.CUdeviceptr(dst), drv.CUdeviceptr(src), nbytes, drv.CUstream(stream_handle)))

    def _tag_ptr(self, tag: int) -> int:
        assert 0 < tag < (1 << 20), "PeerGather: more than 2^20 steps"
        return self.steps.data_ptr() + 4 * tag

    def slot(self):
        b = self.t % self.NBUF
        r0 = self.row0[self.rank]
        return self.full[b][r0:r0 + self.rows[self.rank]]

    def submit(self, producer_stream=None):
        """Pushes the block last handed out by slot() (written on ``producer_stream``, default: the current stream)."""
        import torch
        if producer_stream is None:
            producer_stream = torch.cuda.current_stream(self.device)
        b = self.t % self.NBUF
        tag = self.t + 1
        cs = self.copy_stream.cuda_stream
        drv = self.drv  # noqa: F841
        self.ready.record(producer_stream)
        self.copy_stream.wait_event(self.ready)
        r0, n = self.row0[self.rank], self.rows[self.rank]
        nbytes = n * self.cols * self.esize
        src = self.full[b][r0:r0 + n].data_ptr()
        if self.t >= self.NBUF:
            # every peer has released this matrix (step t - NBUF): its flag word for me holds at least that step's tag
            for q in range(self.world):
                if q != self.rank:
                    self._wait_geq(cs, self.flags[b].data_ptr() + 4 * (self.world + q), tag - self.NBUF)
        for d in range(1, self.world):
            q = (self.rank + d) % self.world     # (every rank starts with another peer: the links are used evenly)
            if nbytes:
                self._copy(self.peer_full[q][b].data_ptr() + r0 * self.cols * self.esize, src, nbytes, cs)
            self._copy(self.peer_flags[q][b].data_ptr() + 4 * self.rank, self._tag_ptr(tag), 4, cs)
        self.t += 1
        return self.full[b]

    def wait(self, stream=None, step=None):
        """Orders ``stream`` behind the arrival of every rank's block of step ``step`` (default: the last one submitted);
        returns the [N, cols] matrix."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        t = self.t - 1 if step is None else step
        b = t % self.NBUF
        for q in range(self.world):
            if q != self.rank:
                self._wait_geq(stream.cuda_stream, self.flags[b].data_ptr() + 4 * q, t + 1)
        return self.full[b]

    def release(self, stream=None, step=None):
        """This rank is done with the matrix of step ``step`` (default: the last one submitted) once ``stream`` has got to
        this point.  The flag copies go through the side stream (a copy-engine operation in the middle of a compute
        stream costs a bubble between its kernels)."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        t = self.t - 1 if step is None else step
        b = t % self.NBUF
        self.done.record(stream)
        self.copy_stream.wait_event(self.done)
        for q in range(self.world):
            if q != self.rank:
                self._copy(self.peer_flags[q][b].data_ptr() + 4 * (self.world + self.rank), self._tag_ptr(t + 1), 4, self.copy_stream.cuda_stream)
