"""One launch of each FeatureTransformer linear shape (for ncu launch lists)."""
import sys
import torch
sys.path.insert(0, ".")
from emip_b200.transformer_layer import linear_tm, layer_norm_tm, mlp_tm   # noqa: E402
L, C = 32 * 1936, 128
x = torch.randn(L, C, device="cuda"); xc = torch.randn(L, 2 * C, device="cuda"); h = torch.randn(L, 8 * C, device="cuda")
w1, w2, w3 = (torch.randn(o, k, device="cuda") * k ** -0.5 for o, k in ((C, C), (8 * C, 2 * C), (C, 8 * C)))
g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
with torch.no_grad():
    for _ in range(2):
        linear_tm(x, w1); linear_tm(xc, w2); linear_tm(h, w3, gelu_in=True); layer_norm_tm(x, g, b, 1e-5, residual=x); mlp_tm(xc, w2, w3)
torch.cuda.synchronize()
