"""PCIe ceiling for the e2e path: pinned 6.2 MB frames up and down concurrently on several streams, no kernels."""
import time
import torch
FB = 6220800
n = 64
hin = [torch.empty(FB, dtype=torch.uint8).pin_memory() for _ in range(n)]
hout = [torch.empty(FB, dtype=torch.uint8).pin_memory() for _ in range(n)]
dev = [torch.empty(FB, dtype=torch.uint8, device="cuda") for _ in range(n)]
for ns in (1, 2, 4, 8, 16):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    def run(reps):
        for r in range(reps):
            for i in range(n):
                with torch.cuda.stream(streams[i % ns]):
                    dev[i].copy_(hin[i], non_blocking=True)
                    hout[i].copy_(dev[i], non_blocking=True)
        torch.cuda.synchronize()
    run(1)
    t0 = time.perf_counter(); run(4); dt = time.perf_counter() - t0
    print(f"{ns:2d} streams: {4 * n / dt:8.0f} frames/s both directions, {4 * n * FB / dt / 1e9:6.1f} GB/s per direction")
# one direction only
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    t0 = time.perf_counter()
    for r in range(4):
        for i in range(n): dev[i].copy_(hin[i], non_blocking=True)
    s.synchronize(); dt = time.perf_counter() - t0
print(f"H2D only: {4 * n * FB / dt / 1e9:.1f} GB/s")
with torch.cuda.stream(s):
    t0 = time.perf_counter()
    for r in range(4):
        for i in range(n): hout[i].copy_(dev[i], non_blocking=True)
    s.synchronize(); dt = time.perf_counter() - t0
print(f"D2H only: {4 * n * FB / dt / 1e9:.1f} GB/s")
