#!/usr/bin/env python
"""Gradient check of the tensor-core backward (a1, a2) against the exact-fp32 CUDA-core backward (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, cases
from emip_b200.matching import global_correlation_softmax as gcs
from emip_b200.flow_attn import FeatureFlowAttention
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
for name, (b, h, w, sc) in {"small": (2, 6, 8, 1.0), "mid": (3, 16, 24, 2.0), "full": (2, 44, 44, 4.1), "lowc": (1, 44, 44, 0.41)}.items():
    f0 = cases.randn(1, (b, 128, h, w), sc).cuda(); f1 = cases.randn(2, (b, 128, h, w), sc).cuda()
    wf = cases.randn(3, (2 * b, 2, h, w)).cuda(); wc = cases.randn(4, (b, h * w, h, w), 0.05).cuda()
    grads = {}
    for exact in (True, False):
        a, c = f0.clone().requires_grad_(True), f1.clone().requires_grad_(True)
        flow, _, corr = gcs(a, c, True, exact_fp32=exact)
        torch.autograd.backward([flow, corr], [wf, wc])
        grads[exact] = (a.grad, c.grad)
    print(f"a1 {name:6s} df0 {rel(grads[False][0], grads[True][0]):.2e} df1 {rel(grads[False][1], grads[True][1]):.2e}")
    a, c = f0.clone().requires_grad_(True), f1.clone().requires_grad_(True)
    flow, _, corr = gcs(a, c, True)
    flow.backward(wf)
    a2, c2 = f0.clone().requires_grad_(True), f1.clone().requires_grad_(True)
    flow, _, corr = gcs(a2, c2, True, exact_fp32=True)
    flow.backward(wf)
    print(f"a1 {name:6s} (dflow only) df0 {rel(a.grad, a2.grad):.2e} df1 {rel(c.grad, c2.grad):.2e}")
    m = FeatureFlowAttention(128).cuda()
    x = torch.cat((f0, f1), 0)
    fl = cases.randn(5, (2 * b, 2, h, w), 8.0).cuda(); wo = cases.randn(6, (2 * b, 2, h, w)).cuda()
    g = {}
    for exact in (True, False):
        m.exact_fp32 = exact
        xx = x.clone().requires_grad_(True)
        m(xx, fl).backward(wo)
        g[exact] = xx.grad
        m.zero_grad()
    print(f"a2 {name:6s} dx {rel(g[False], g[True]):.2e}")
