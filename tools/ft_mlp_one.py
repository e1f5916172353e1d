"""One fused MLP call (two gemm_tc_kernel launches: 256 -> 1024 with the GELU / split epilogue, 1024 -> 128) for ncu --set full."""
import sys
import torch
sys.path.insert(0, ".")
from emip_b200.transformer_layer import mlp_tm   # noqa: E402
L, C = 32 * 1936, 128
xc = torch.randn(L, 2 * C, device="cuda")
w2, w3 = (torch.randn(o, k, device="cuda") * k ** -0.5 for o, k in ((8 * C, 2 * C), (C, 8 * C)))
with torch.no_grad():
    mlp_tm(xc, w2, w3)
torch.cuda.synchronize()
