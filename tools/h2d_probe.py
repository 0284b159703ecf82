#!/usr/bin/env python3
"""Developer probe: pinned host -> device copy of one step's chain state (4 x 31 x 10000 int8), as one
copy and split over several streams."""
import time
import torch
dev = torch.device("cuda:0")
nbytes = 1_240_000
h = torch.zeros(nbytes, dtype=torch.int8).pin_memory()
d = torch.zeros(nbytes, dtype=torch.int8, device=dev)
for k in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    chunk = nbytes // k
    def go():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
    for _ in range(5):
        go()
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        go()
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"{k} stream(s): {dt * 1e6:.1f} us per {nbytes} B -> {nbytes / dt / 1e9:.1f} GB/s")
