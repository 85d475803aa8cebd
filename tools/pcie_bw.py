#!/usr/bin/env python
"""Host <-> device copy bandwidth of this box with pinned memory (the bound of bench.py's `e2e`): one JSON line."""
import json
import torch

n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=8):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


h2d = timed(lambda: d.copy_(h, non_blocking=True))
d2h = timed(lambda: h.copy_(d, non_blocking=True))
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


bi = timed(both)
print(json.dumps({"pinned_h2d_GBps": round(h2d, 1), "pinned_d2h_GBps": round(d2h, 1), "bidirectional_each_GBps": round(bi, 1),
                  "bytes_per_copy": n}))
