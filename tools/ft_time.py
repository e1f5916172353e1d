#!/usr/bin/env python
"""FeatureTransformer call time under the gemm_tc diagnostics switches (emip_debug_gemm_wide_tiles: bit 0 = 256-column tiles for
the 256 -> 1024 layer): ft_time.py [pairs] [flags ...]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from emip_b200 import _lib
from emip_b200 import chain as ch

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
torch.manual_seed(1)
m = ch._FeatureTransformer().to(dev)
x = torch.randn(2 * pairs, 1936, 128, device=dev)
L = _lib.lib()
ref = None
flag_list = [int(a) for a in sys.argv[2:]] or [0, 1]
for wide in flag_list * 2:
    L.emip_debug_gemm_wide_tiles(wide)
    with torch.no_grad():
        for _ in range(2):
            y = ch.feature_transformer_tokens(x, m, 44, 44, 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = ch.feature_transformer_tokens(x, m, 44, 44, 2)
        e1.record()
        torch.cuda.synchronize()
    if ref is None:
        ref = y.clone()
    print(f"gemm flags {wide}: {e0.elapsed_time(e1) / 5:7.3f} ms per FeatureTransformer call   max|diff| {(y - ref).abs().max().item():.1e}")
L.emip_debug_gemm_wide_tiles(0)
