"""One f2-shaped attention call (128 windows of 484 tokens) for ncu."""
import sys, torch
sys.path.insert(0, ".")
from emip_b200.window_attn import attention
nb, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 484)
g = torch.Generator().manual_seed(5)
q, k, v = (torch.randn(nb, n, 128, generator=g).cuda() for _ in range(3))
with torch.no_grad():
    for _ in range(3):
        out = attention(q, k, v)
torch.cuda.synchronize()
print(out.float().abs().mean().item())
