"""a2 (FeatureFlowAttention incl. its two projections) forward / forward+backward at 2B = 32."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
from emip_b200.flow_attn import FeatureFlowAttention
g = torch.Generator(device="cuda").manual_seed(0)
m = FeatureFlowAttention(128).cuda()
x = 4.1 * torch.randn(32, 128, 44, 44, device="cuda", generator=g)
fl = 8 * torch.randn(32, 2, 44, 44, device="cuda", generator=g)
wo = torch.randn(32, 2, 44, 44, device="cuda", generator=g)


def t(fn, it=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it


def f():
    with torch.no_grad(): m(x, fl)


def fb():
    xx = x.detach().requires_grad_(True)
    m(xx, fl).backward(wo)
    m.zero_grad(set_to_none=True)


print("a2 fwd %.3f ms, fwd+bwd %.3f ms (all parameters require grad)" % (t(f), t(fb)))
for p in m.parameters(): p.requires_grad_(False)
print("a2 with frozen weights (as the reference trains): fwd+bwd %.3f ms" % t(fb))
m.exact_fp32 = True
print("exact-fp32 path (library projections + CUDA-core kernels): fwd %.3f ms" % t(f, 3))
