"""One attention forward + backward at the f2 window shape (128 problems x 484 tokens) for ncu."""
import sys, torch
sys.path.insert(0, ".")
from emip_b200.window_attn import attention
g = torch.Generator().manual_seed(5)
q, k, v, w = (torch.randn(128, 484, 128, generator=g).cuda() for _ in range(4))
for _ in range(2):
    a, b, c = (t.detach().requires_grad_(True) for t in (q, k, v))
    attention(a, b, c).backward(w)
torch.cuda.synchronize()
