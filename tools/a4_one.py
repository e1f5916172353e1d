"""One Injector forward + backward (B = 16) for ncu."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
import cases
from emip_b200.injector import Injector
B = 16
g = torch.Generator(device="cuda").manual_seed(0)
inj = Injector().cuda(); inj.transformer.load_state_dict(cases.injector_params(7))
x = 2.2 * torch.randn(B, 128, 44, 44, device="cuda", generator=g); y = torch.randn(B, 128, 44, 44, device="cuda", generator=g)
w = torch.randn(B, 128, 44, 44, device="cuda", generator=g)
for _ in range(2):
    a, b = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    inj(a, b).backward(w)
torch.cuda.synchronize()
