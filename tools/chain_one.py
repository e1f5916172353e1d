#!/usr/bin/env python
"""A few passes of the chained hot path for ncu launch lists / captures: chain_one.py [pairs] [passes]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from emip_b200.chain import MotionChain  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(123)
dev = torch.device("cuda", 0)
m = MotionChain().to(dev).eval()
gm = 2.2 * torch.randn(2 * pairs, 128, 44, 44, device=dev)
seg = torch.randn(2 * pairs, 128, 44, 44, device=dev)
with torch.no_grad():
    m(gm, seg)                                   # warm-up: prepared weights, function attributes (not profiled with --profile-from-start off)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(passes - 1):
        m(gm, seg)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done", pairs, passes)
