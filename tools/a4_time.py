#!/usr/bin/env python
"""Kernel-level timing of the Injector (a4) forward / backward (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, cases
from torch.profiler import profile, ProfilerActivity
from emip_b200.injector import Injector
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator(device="cuda").manual_seed(0)
inj = Injector().cuda(); inj.transformer.load_state_dict(cases.injector_params(7))
x = 2.2 * torch.randn(B, 128, 44, 44, device="cuda", generator=g); y = torch.randn(B, 128, 44, 44, device="cuda", generator=g)
w = torch.randn(B, 128, 44, 44, device="cuda", generator=g)
def fwd():
    with torch.no_grad(): inj(x, y)
def fb():
    a, b = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    inj(a, b).backward(w); inj.zero_grad(set_to_none=True)
for name, fn in (("fwd", fwd), ("fwd+bwd", fb)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): fn()
        torch.cuda.synchronize()
    print("==", name, "B =", B)
    tot = 0
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]:
        print(f"   {e.key[:60]:60s} n={e.count // 3:3d}/call  {e.device_time_total / 3:9.1f} us/call")
    print("   total", sum(e.device_time_total for e in prof.key_averages()) / 3)
