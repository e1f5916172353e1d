#!/usr/bin/env python
"""Per-role wait / work cycles of mlp_fused_kernel inside a FeatureTransformer call (diagnostic; run on the GPU box):
mlp_roles.py [pairs].  The counters of the six launches of the call add up (the kernel writes, the last launch wins: one block)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from emip_b200 import _lib
from emip_b200 import chain as ch

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
torch.manual_seed(1)
m = ch._FeatureTransformer()
m.layers = m.layers[:1]
m = m.to(dev)
x = torch.randn(2 * pairs, 1936, 128, device=dev)
L = _lib.lib()
names = ["math: wait s_full", "math: tmem ld + GELU + split", "math: wait h_empty", "math: tmem st + arrive", "math: wait o_full", "math: LayerNorm epilogue",
         "math: total", "mma: wait s_empty", "mma: wait ring", "mma: wait h_full", "mma: wait o_empty", "mma: wait x_full", "mma: total",
         "tma: wait ring slot / x_empty", "mma: wait ring (UMMA-2 part of it)", "tma: total"]
prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
with torch.no_grad():
    for rep in range(3):
        L.emip_attn_tc_set_profile_buffer(ctypes.c_void_p(prof.data_ptr() if rep == 2 else 0))
        ch.feature_transformer_tokens(x, m, 44, 44, 2)
    torch.cuda.synchronize()
    L.emip_attn_tc_set_profile_buffer(ctypes.c_void_p(0))
p = prof.view(148, 16).double().cpu()
tiles = 2 * pairs * 1936 // 128
per = -(-tiles // 148)
full = p[: tiles % 148 or 148].mean(0)
print(f"== {tiles} row tiles; CTAs with {per} tiles x 8 hidden blocks (mean cycles over those CTAs)")
for i, nm in enumerate(names):
    print(f"   {nm:32s} {full[i]:10.0f}   per hidden block {full[i] / (per * 8):8.0f}")
