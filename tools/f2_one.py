"""One plain + one shifted split-window attention layer (2B = 32) for ncu."""
import sys, torch
sys.path.insert(0, ".")
from emip_b200.window_attn import single_head_split_window_attention as swa
g = torch.Generator().manual_seed(3)
q, k, v = (torch.randn(32, 44 * 44, 128, generator=g).cuda() for _ in range(3))
m = torch.zeros(1, device="cuda")
with torch.no_grad():
    for _ in range(2):
        a = swa(q, k, v, 2, False, 44, 44, None)
        b = swa(q, k, v, 2, True, 44, 44, m)
torch.cuda.synchronize()
