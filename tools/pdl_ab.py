#!/usr/bin/env python
"""Same-box A / B of programmatic dependent launch: the CUDA-graph chain step with and without it (GraphedChain(pdl=...) =
emip_set_programmatic_launch around the capture; the graph is re-captured for each setting): pdl_ab.py [pairs ...]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from emip_b200 import _lib
from emip_b200.chain import MotionChain, GraphedChain

L = _lib.lib()
dev = torch.device("cuda", 0)
for pairs in [int(a) for a in sys.argv[1:]] or [64, 8]:
    torch.manual_seed(123)
    m = MotionChain().to(dev).eval()
    gm = 2.2 * torch.randn(2 * pairs, 128, 44, 44, device=dev)
    seg = torch.randn(2 * pairs, 128, 44, 44, device=dev)
    res = {}
    for rnd in range(3):
        for flag in (0, 16):
            g = GraphedChain(m, gm, seg, pdl=(flag == 0))
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            res.setdefault(flag, []).append(e0.elapsed_time(e1) / n)
            del g
    print(f"{pairs} pairs: with PDL {min(res[0]):.4f} ms {['%.3f' % v for v in res[0]]} | without {min(res[16]):.4f} ms {['%.3f' % v for v in res[16]]}")
