#!/usr/bin/env python
"""One flow field through flow_warp forward + backward (for ncu captures): k3_one.py <variant> <field> [iters]."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import k3_bench
from emip_b200 import _lib
from emip_b200._lib import I, LL, ptr

variant, field = int(sys.argv[1]), sys.argv[2]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
L = _lib.lib()
kern = k3_bench.variant_kernel(variant) if variant else 0
nul = ctypes.c_void_p(None)
g = torch.Generator(device=dev).manual_seed(5)
B, C, H, W = k3_bench.B, k3_bench.C, k3_bench.H, k3_bench.W
x = torch.randn(B, C, H, W, device=dev, generator=g)
out = torch.empty_like(x)
dout = torch.randn(B, C, H, W, device=dev, generator=g)
dflow = torch.empty(B, 2, H, W, device=dev)
f = k3_bench.flows(dev, g)[field][:, 2:]
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(iters):
    _lib.check(L.emip_flow_warp_fwd_ex(ptr(x), ptr(f), ptr(out), I(B), I(C), I(H), I(W), LL(f.stride(0)), LL(f.stride(1)), I(0), I(kern), nul, nul, ctypes.c_uint(0), sp), "fwd")
    _lib.check(L.emip_flow_warp_bwd_ex(ptr(x), ptr(f), ptr(dout), ptr(dflow), None, I(B), I(C), I(H), I(W), LL(f.stride(0)), LL(f.stride(1)), I(0), I(kern), nul, nul, ctypes.c_uint(0), sp), "bwd")
torch.cuda.synchronize()
print("done", variant, field)
