#!/usr/bin/env python
"""Timing breakdown of the a1 / a2 backward paths (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
from torch.profiler import profile, ProfilerActivity
from emip_b200.matching import global_correlation_softmax as gcs
from emip_b200.flow_attn import FeatureFlowAttention
B, C, H, W = 16, 128, 44, 44
g = torch.Generator(device="cuda").manual_seed(0)
f0 = 4.1 * torch.randn(B, C, H, W, device="cuda", generator=g); f1 = 4.1 * torch.randn(B, C, H, W, device="cuda", generator=g)
wf = torch.randn(2 * B, 2, H, W, device="cuda", generator=g); wc = 0.05 * torch.randn(B, H * W, H, W, device="cuda", generator=g)
m = FeatureFlowAttention(128).cuda()
x = torch.cat((f0, f1), 0); fl = 8 * torch.randn(2 * B, 2, H, W, device="cuda", generator=g); wo = torch.randn(2 * B, 2, H, W, device="cuda", generator=g)
def run(which):
    a, c = f0.clone().requires_grad_(True), f1.clone().requires_grad_(True)
    if which == "a2":
        xx = x.clone().requires_grad_(True); m(xx, fl).backward(wo); return
    flow, _, corr = gcs(a, c, True)
    if which == "both": torch.autograd.backward([flow, corr], [wf, wc])
    elif which == "flow": flow.backward(wf)
    else: corr.backward(wc)
for which in ("both", "flow", "corr", "a2"):
    for _ in range(2): run(which)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): run(which)
        torch.cuda.synchronize()
    print("==", which)
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:7]
    for e in rows:
        print(f"   {e.key[:70]:70s} n={e.count:3d} avg {e.device_time_total / e.count:9.1f} us")
