#!/usr/bin/env python
"""a1 forward (tcgen05 path) under the two work schedules of match_tc_fwd, by batch size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from emip_b200 import _lib
from emip_b200.matching import global_correlation_softmax

L = _lib.lib()
for B in (1, 2, 4, 8, 10, 16, 24):
    f0 = 4.1 * torch.randn(B, 128, 44, 44, device="cuda")
    f1 = 4.1 * torch.randn(B, 128, 44, 44, device="cuda")
    row = []
    for sched in ("items", "stream_k", None):
        for _ in range(5):
            global_correlation_softmax(f0, f1, True, schedule=sched)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            global_correlation_softmax(f0, f1, True, schedule=sched)
        e1.record()
        torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 30 * 1e3)
    print(f"B={B:3d}  strided items {row[0]:7.1f} us   stream-K {row[1]:7.1f} us   auto {row[2]:7.1f} us   (split + fused launch, corr emitted)")
