"""One conv_corr[0] forward + backward at B = 16 for ncu."""
import sys, torch
sys.path.insert(0, ".")
from emip_b200.conv_corr import conv_corr_first_layer
B, C, H, W, O = 16, 128, 44, 44, 968
g = torch.Generator(device="cuda").manual_seed(0)
f0 = 4.1 * torch.randn(B, C, H, W, device="cuda", generator=g); f1 = 4.1 * torch.randn(B, C, H, W, device="cuda", generator=g)
w = (torch.randn(O, H * W, 3, 3, device="cuda", generator=g) * 0.01).requires_grad_(True)
b = torch.zeros(O, device="cuda", requires_grad=True)
wo = torch.randn(B, O, H, W, device="cuda", generator=g)
for _ in range(2):
    a, c = f0.detach().requires_grad_(True), f1.detach().requires_grad_(True)
    conv_corr_first_layer(a, c, w, b).backward(wo)
torch.cuda.synchronize()
