#!/usr/bin/env python
"""Convex x8 upsampling (f4) micro-benchmark: achieved HBM GB/s forward / backward at 2B = 32, 44x44."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from emip_b200.upsample import upsample_flow_convex
B, h, w = 32, 44, 44
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6542.1
g = torch.Generator(device="cuda").manual_seed(1)
flows = [5 * torch.randn(B, 2, h, w, device="cuda", generator=g) for _ in range(4)]
masks = [torch.randn(B, 576, h, w, device="cuda", generator=g) for _ in range(4)]     # 4 x 143 MB > L2
wout = torch.randn(B, 2, 8 * h, 8 * w, device="cuda", generator=g)
def t(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
i = [0]
def fwd():
    i[0] += 1
    with torch.no_grad(): upsample_flow_convex(flows[i[0] % 4], masks[i[0] % 4])
def fb():
    i[0] += 1
    f, m = flows[i[0] % 4].detach().requires_grad_(True), masks[i[0] % 4].detach().requires_grad_(True)
    upsample_flow_convex(f, m).backward(wout)
fb_bytes = B * h * w * 4 * (578 + 128)
bb_bytes = B * h * w * 4 * (578 + 128 + 578)
mf = t(fwd); mfb = t(fb, 10)
print(f"f4 convex upsample fwd {mf*1e3:.1f} us  {fb_bytes/mf/1e6:.0f} GB/s ({100*fb_bytes/mf/1e6/pk:.1f}% of {pk:.0f})   "
      f"bwd {(mfb-mf)*1e3:.1f} us  {bb_bytes/(mfb-mf)/1e6:.0f} GB/s ({100*bb_bytes/(mfb-mf)/1e6/pk:.1f}%)")
