#!/usr/bin/env python
"""flow_warp (K3) micro-benchmark: achieved HBM GB/s, forward and backward-to-flow, three flow fields.

    model_like : per-sample translation (sigma 10 px) + smooth x8-upsampled variation (sigma 0.5 px per 8-px cell,
                 local flow gradient ~0.07 px/px = a few percent of zoom / rotation / deformation), i.e. what a
                 trained GMFlow emits after convex upsampling (piecewise smooth)
    warpy      : same with sigma 2 px per cell (gradient ~0.35 px/px: a 35 % local stretch; upper end of plausible)
    iid5px / iid20px : SURVEY.md 8(d) stress cases (independent per-pixel flow; gather-bound)
"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import cases
from emip_b200 import _lib
from emip_b200._lib import I, LL, ptr

B, C, H, W = 64, 3, 352, 352


def flows(dev, g):
    base = 10.0 * torch.randn(B, 2, 1, 1, device=dev, generator=g)
    def smooth(mag):
        sm = torch.cat([cases.smooth_flow(11, B, H, W, mag), cases.smooth_flow(12, B, H, W, mag)], 1).to(dev)
        sm[:, :2] += base
        sm[:, 2:] -= base
        return sm
    return {"model_like": smooth(0.5), "warpy": smooth(2.0), "iid5px": 5.0 * torch.randn(B, 4, H, W, device=dev, generator=g),
            "iid20px": 20.0 * torch.randn(B, 4, H, W, device=dev, generator=g)}


def variant_kernel(v):
    """1-99: direct gather, v CTAs per SM; 1xx: staged (auto), v - 100 CTAs per SM; 2xx: staged only."""
    if v >= 200:
        return 2 | ((v - 200) << 8)
    if v >= 100:
        return 0 | ((v - 100) << 8)
    return 1 | (v << 8)


def run(dev, hbm_peak, iters=50, variant=0):
    from emip_b200 import warp
    L = _lib.lib()
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(B, C, H, W, device=dev, generator=g)
    out = torch.empty_like(x)
    dout = torch.randn(B, C, H, W, device=dev, generator=g)
    dflow = torch.empty(B, 2, H, W, device=dev)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    fw_bytes = B * H * W * (2 * C + 2) * 4
    bw_bytes = B * H * W * (2 * C + 2 + 2) * 4
    res = {}
    for name, fl in flows(dev, g).items():
        f = fl[:, 2:]      # the channel-slice call pattern of loss_flow.py:91

        def pick():
            if variant == 0:
                return warp._choice(dev).pick()           # the product's adaptive choice (emip_b200/warp.py)
            return variant_kernel(variant), None, None, 0

        def fwd():
            k, ds, hs, seq = pick()
            _lib.check(L.emip_flow_warp_fwd_ex(ptr(x), ptr(f), ptr(out), I(B), I(C), I(H), I(W), LL(f.stride(0)), LL(f.stride(1)), I(0),
                                               I(k), ctypes.c_void_p(ds), ctypes.c_void_p(hs), ctypes.c_uint(seq), sp), "fwd")

        def bwd():
            k, ds, hs, seq = pick()
            _lib.check(L.emip_flow_warp_bwd_ex(ptr(x), ptr(f), ptr(dout), ptr(dflow), None, I(B), I(C), I(H), I(W), LL(f.stride(0)),
                                               LL(f.stride(1)), I(0), I(k), ctypes.c_void_p(ds), ctypes.c_void_p(hs), ctypes.c_uint(seq),
                                               sp), "bwd")
        r = {}
        for tag, fn, nbytes in (("fwd", fwd, fw_bytes), ("bwd", bwd, bw_bytes)):
            for _ in range(10):
                fn()
            torch.cuda.synchronize()      # lets the kernel-choice feedback land (emip_b200/warp.py)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            r[tag] = {"launch_ms": ms, "achieved": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak,
                      "bytes_per_launch": nbytes}
        res[name] = r
    return res


if __name__ == "__main__":
    pk = 6542.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        pk = json.load(open(p))["hbm_gbs"]
    import sys as _sys
    variants = [int(a, 0) for a in _sys.argv[1:]] or [0]
    for variant in variants:
      print("variant (0 = product: adaptive choice in emip_b200/warp.py; 1-99 direct, 1xx staged if covered, 2xx staged only):", variant)
      r = run(torch.device("cuda", 0), pk, variant=variant)
      for k, v in r.items():
          print(f"{k:11s} fwd {v['fwd']['launch_ms']*1e3:7.1f} us {v['fwd']['achieved']:7.0f} GB/s ({100*v['fwd']['frac']:4.1f}%)   "
                f"bwd {v['bwd']['launch_ms']*1e3:7.1f} us {v['bwd']['achieved']:7.0f} GB/s ({100*v['bwd']['frac']:4.1f}%)")
