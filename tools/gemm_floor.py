#!/usr/bin/env python
"""Where the time of the FeatureTransformer GEMMs goes: each token-row layer timed as it ships and with the epilogue warps
idle (emip_debug_gemm_wide_tiles bit 1: only TMA operand streaming + UMMA remain; NOTE that this also removes the epilogue's HBM
traffic) and with the operand loads off as well (bit 3: UMMA issue alone).  Each call includes the fp32 -> bf16 hi | lo operand
split kernel of its input.  gemm_floor.py [pairs]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from emip_b200 import _lib
from emip_b200.transformer_layer import linear_tm, linear_ln_tm, mlp_tm, linear_tm_multi

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
L = _lib.lib()
rows = 2 * pairs * 1936
x = torch.randn(rows, 128, device=dev)
x2 = torch.randn(rows, 256, device=dev)
w = torch.randn(128, 128, device=dev) / 11
w1 = torch.randn(1024, 256, device=dev) / 16
w2 = torch.randn(128, 1024, device=dev) / 32
g, b = torch.ones(128, device=dev), torch.zeros(128, device=dev)
cases = {
    "linear 128->128 (fp32 out)": lambda: linear_tm(x, w),
    "3 x linear 128->128 on one operand split (q | k | v)": lambda: linear_tm_multi(x, [w, w, w]),
    "linear 128->128 + LN + residual": lambda: linear_ln_tm(x, w, g, b, 1e-5, residual=x),
    "mlp 256->1024 (GELU, hi | lo out) ->128 + LN + residual": lambda: mlp_tm(x2, w1, w2, g, b, 1e-5, residual=x),
}


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


with torch.no_grad():
    for name, fn in cases.items():
        res = {}
        for flags in (0, 1, 2, 10, 11) * 2:
            L.emip_debug_gemm_wide_tiles(flags)
            res[flags] = min(res.get(flags, 1e9), timed(fn))
        L.emip_debug_gemm_wide_tiles(0)
        print(f"{name}: {res[0]:7.1f} us per call (256-column tiles for N >= 512: {res[1]:7.1f}) | epilogue warps idle: {res[2]:7.1f} us | "
              f"no epilogue, no TMA loads (UMMA issue alone): {res[10]:7.1f} us, with 256-column tiles {res[11]:7.1f} us  ({rows} rows)")
