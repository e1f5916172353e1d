"""One FeatureTransformer feed-forward call (mlp + norm2 + residual: the operand splits + mlp_fused_kernel) for ncu --set full:
ft_mlp_one.py [maps]."""
import sys
import torch
sys.path.insert(0, ".")
from emip_b200.transformer_layer import mlp_tm   # noqa: E402
maps = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L, C = maps * 1936, 128
xc = torch.randn(L, 2 * C, device="cuda")
res = torch.randn(L, C, device="cuda")
w2, w3 = (torch.randn(o, k, device="cuda") * k ** -0.5 for o, k in ((8 * C, 2 * C), (C, 8 * C)))
g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
with torch.no_grad():
    mlp_tm(xc, w2, w3, g, b, 1e-5, residual=res)
torch.cuda.synchronize()
