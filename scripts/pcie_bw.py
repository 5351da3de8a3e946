"""Pinned host<->device copy bandwidth of the box (context for the e2e number)."""
import time

import torch

n = 1 << 30
h1, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d1, d2 = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9


def both():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


print("H2D %.1f GB/s  D2H %.1f GB/s  both directions at once %.1f GB/s each" %
      (t(lambda: d1.copy_(h1, non_blocking=True)), t(lambda: h2.copy_(d2, non_blocking=True)), t(both)))
