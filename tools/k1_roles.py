#!/usr/bin/env python
"""Per-role wait-cycle breakdown of the tcgen05 matching kernel (diagnostic; run on the GPU box)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from emip_b200 import _lib
from emip_b200._lib import I, SZ, ptr
from emip_b200._ws import workspace

B, C, H, W = 16, 128, 44, 44
N = H * W
dev = torch.device("cuda", 0)
L = _lib.lib()
g = torch.Generator(device=dev).manual_seed(1)
f0 = 4.1 * torch.randn(B, C, H, W, device=dev, generator=g)
f1 = 4.1 * torch.randn(B, C, H, W, device=dev, generator=g)
L.emip_global_matching_workspace.restype = ctypes.c_size_t
ws, wp, wn = workspace(L.emip_global_matching_workspace(I(B), I(C), I(H), I(W)), dev)
flow = torch.empty(2 * B, 2, H, W, device=dev)
corr = torch.empty(B, N, H, W, device=dev)
prof = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
names = ["prod:q_empty", "prod:k_empty", "mma:s_empty", "mma:k_full", "mma:q_full", "mma:total", "smx:s_full", "smx:total"]
import itertools
for ts, (label, flags, ct) in itertools.product((8, 16), (("3-term +corr", 0, corr), ("3-term flow", 0, None), ("bf16 +corr", 4, corr), ("bf16 flow", 4, None))):
    L.emip_match_tc_set_variant(ts)
    label = f"{ts} softmax warps, " + label
    for rep in range(3):
        L.emip_match_tc_set_profile_buffer(ctypes.c_void_p(prof.data_ptr() if rep == 2 else 0))
        _lib.check(L.emip_global_matching_fwd(ptr(f0), ptr(f1), ptr(flow), ptr(ct), None, ctypes.c_void_p(wp), SZ(wn),
                                              I(B), I(C), I(H), I(W), I(1), I(flags | (2 if rep else 0)), sp), "fwd")
    torch.cuda.synchronize()
    L.emip_match_tc_set_profile_buffer(ctypes.c_void_p(0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        _lib.check(L.emip_global_matching_fwd(ptr(f0), ptr(f1), ptr(flow), ptr(ct), None, ctypes.c_void_p(wp), SZ(wn),
                                              I(B), I(C), I(H), I(W), I(1), I(flags | 2), sp), "fwd")
    e1.record()
    torch.cuda.synchronize()
    label += f"  [{e0.elapsed_time(e1) / 20 * 1e3:.1f} us/launch]"
    p = prof.view(148, 8).double().cpu()
    two = p[:108].mean(0)   # CTAs with two work items
    one = p[108:].mean(0)
    print(f"== {label}: mean cycles over CTAs with 2 items | 1 item")
    for i, n in enumerate(names):
        print(f"   {n:14s} {two[i]:12.0f} | {one[i]:12.0f}")
L.emip_match_tc_set_variant(0)
