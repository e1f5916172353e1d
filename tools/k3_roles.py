#!/usr/bin/env python
"""Wait-cycle profile of the staged flow_warp kernel (producer / worker roles): k3_roles.py [field]."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import k3_bench
from emip_b200 import _lib
from emip_b200._lib import I, LL, ptr
field = sys.argv[1] if len(sys.argv) > 1 else "model_like"
dev = torch.device("cuda", 0)
L = _lib.lib()
kern = k3_bench.variant_kernel(202)
nul = ctypes.c_void_p(None)
g = torch.Generator(device=dev).manual_seed(5)
B, C, H, W = k3_bench.B, k3_bench.C, k3_bench.H, k3_bench.W
x = torch.randn(B, C, H, W, device=dev, generator=g); out = torch.empty_like(x)
dout = torch.randn(B, C, H, W, device=dev, generator=g); dflow = torch.empty(B, 2, H, W, device=dev)
f = k3_bench.flows(dev, g)[field][:, 2:]
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
prof = torch.zeros(296, 8, dtype=torch.int64, device=dev)
def fwd(): _lib.check(L.emip_flow_warp_fwd_ex(ptr(x), ptr(f), ptr(out), I(B), I(C), I(H), I(W), LL(f.stride(0)), LL(f.stride(1)), I(0), I(kern), nul, nul, ctypes.c_uint(0), sp), "fwd")
def bwd(): _lib.check(L.emip_flow_warp_bwd_ex(ptr(x), ptr(f), ptr(dout), ptr(dflow), None, I(B), I(C), I(H), I(W), LL(f.stride(0)), LL(f.stride(1)), I(0), I(kern), nul, nul, ctypes.c_uint(0), sp), "bwd")
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    for _ in range(3): fn()
    L.emip_debug_flow_warp_staged_profile(ctypes.c_void_p(prof.data_ptr()))
    prof.zero_(); fn(); torch.cuda.synchronize()
    L.emip_debug_flow_warp_staged_profile(None)
    p = prof.double().mean(0).tolist()
    print(f"{name} {field}: producer total {p[0]:.0f} cyc: wait flow_empty {p[1]:.0f} box_full {p[2]:.0f} win_empty {p[3]:.0f} | "
          f"worker total {p[4]:.0f}: wait flow_full {p[5]:.0f} win_full {p[6]:.0f} tiles {p[7]:.1f}")
