#!/usr/bin/env python
"""Numerical check of the tcgen05 matching kernel variants against the exact-fp32 CUDA-core kernel (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, cases
from emip_b200 import _lib
from emip_b200.matching import global_correlation_softmax as gcs
L = _lib.lib()
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
for name, (b, h, w, sc) in {"small": (2, 6, 8, 1.0), "ragged": (1, 13, 11, 2.0), "mid": (2, 16, 16, 2.0), "full": (2, 44, 44, 4.1)}.items():
    f0 = cases.randn(1, (b, 128, h, w), sc).cuda(); f1 = cases.randn(2, (b, 128, h, w), sc).cuda()
    fe, _, ce = gcs(f0, f1, True, exact_fp32=True)
    for ts in (8, 16):
        L.emip_match_tc_set_variant(ts)
        for bf in (False, True):
            f, _, c = gcs(f0, f1, True, bf16=bf)
            cc = c.view(b, h * w, h * w)
            print(f"{name:7s} ts={ts} bf16={int(bf)}  flow {rel(f, fe):.2e}  corr {rel(c, ce):.2e}  corr[:, :, :h*w//2] {rel(cc[:, :, :h*w//2], ce.view(b,h*w,h*w)[:, :, :h*w//2]):.2e}")
L.emip_match_tc_set_variant(0)
