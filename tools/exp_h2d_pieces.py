"""Host-to-device copy rate of one pinned 5 GB buffer: one copy, 1,000 copies of ~5 MB (one stream, two, four), odd source offsets."""
import time, torch
N = 1000; SZ = 5063061
host = torch.empty(N * SZ + 4096, dtype=torch.uint8, pin_memory=True); host.zero_()
dev = torch.empty(N * 5063168 + 4096, dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]
def run(ns, piece, src_shift=0, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if piece:
            for i in range(N):
                with torch.cuda.stream(streams[i % ns]):
                    dev[i * 5063168:i * 5063168 + SZ].copy_(host[src_shift + i * SZ:src_shift + (i + 1) * SZ], non_blocking=True)
        else:
            with torch.cuda.stream(streams[0]):
                dev[:N * SZ].copy_(host[:N * SZ], non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return N * SZ / best / 1e9
print("one copy            %.1f GB/s" % run(1, False))
for ns in (1, 2, 4):
    print("1000 copies, %d streams %.1f GB/s" % (ns, run(ns, True)))
print("1000 copies, 1 stream, source offset +1  %.1f GB/s" % run(1, True, 1))
print("1000 copies, 2 streams, source offset +1 %.1f GB/s" % run(2, True, 1))
