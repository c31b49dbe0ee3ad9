#!/usr/bin/env python
"""Pinned host->device copy bandwidth of the box (what bounds the e2e arm: 1.9 MB per sweep must cross PCIe)."""
import json, torch
n = 1 << 28
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for chunks in (1, 128):
    s = torch.cuda.Stream()
    step = n // chunks
    with torch.cuda.stream(s):
        for _ in range(2):
            for c in range(chunks):
                d[c * step:(c + 1) * step].copy_(h[c * step:(c + 1) * step], non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(4):
            for c in range(chunks):
                d[c * step:(c + 1) * step].copy_(h[c * step:(c + 1) * step], non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    out[f"h2d_gbs_{chunks}_chunks"] = 4 * n / 1e9 / (e0.elapsed_time(e1) * 1e-3)
print(json.dumps(out))
