"""Forward + backward time of the re-associated conv_corr[0] (backward = library ops today) vs cuDNN on the materialised volume."""
import sys, math, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from emip_b200.conv_corr import conv_corr_first_layer
B, C, H, W, O = 16, 128, 44, 44, 968
g = torch.Generator(device="cuda").manual_seed(0)
f0 = 4.1 * torch.randn(B, C, H, W, device="cuda", generator=g); f1 = 4.1 * torch.randn(B, C, H, W, device="cuda", generator=g)
w = (torch.randn(O, H * W, 3, 3, device="cuda", generator=g) * 0.01).requires_grad_(True)
b = torch.zeros(O, device="cuda", requires_grad=True)
wo = torch.randn(B, O, H, W, device="cuda", generator=g)


def timeit(fn, iters=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def ours():
    a, c = f0.detach().requires_grad_(True), f1.detach().requires_grad_(True)
    conv_corr_first_layer(a, c, w, b).backward(wo)
    w.grad = None; b.grad = None


def ref():
    a, c = f0.detach().requires_grad_(True), f1.detach().requires_grad_(True)
    corr = torch.matmul(a.view(B, C, -1).permute(0, 2, 1), c.view(B, C, -1)).view(B, H, W, H * W).permute(0, 3, 1, 2) / math.sqrt(C)
    F.conv2d(corr, w, b, padding=1).backward(wo)
    w.grad = None; b.grad = None


def fwd():
    with torch.no_grad(): conv_corr_first_layer(f0, f1, w, b)


print(f"ours fwd {timeit(fwd):.3f} ms, fwd+bwd {timeit(ours):.3f} ms; reference composition (matmul + cuDNN conv on corr, TF32 conv) fwd+bwd {timeit(ref):.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2): ours()
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:10]:
    print(f"   {e.key[:80]:80s} n={e.count // 2:3d}/call {e.device_time_total / 2:9.1f} us/call")
