"""One photometric loss term forward + backward (B = 64, 3 x 352 x 352) for ncu."""
import sys, torch
sys.path.insert(0, ".")
from emip_b200.photometric import photometric_loss
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(64, 3, 352, 352, device="cuda", generator=g)
rec = x + 0.2 * torch.randn(64, 3, 352, 352, device="cuda", generator=g)
occ = (torch.rand(64, 1, 352, 352, device="cuda", generator=g) > 0.2).float()
for _ in range(2):
    r = rec.detach().requires_grad_(True)
    photometric_loss(x, r, occ).backward()
torch.cuda.synchronize()
