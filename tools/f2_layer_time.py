"""Whole-layer timing of the split-window attention (2B = 32 feature maps of 44x44 tokens), plain and shifted."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
from emip_b200.window_attn import single_head_split_window_attention as swa
from oracle import restate as O
H = W = 44
g = torch.Generator().manual_seed(3)
q, k, v = (torch.randn(32, H * W, 128, generator=g).cuda() for _ in range(3))
amask = torch.zeros(1, device="cuda")


def timeit(fn, iters=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


wg = torch.randn(32, H * W, 128, generator=g).cuda()
for shift in (False, True):
    def fb():
        a, b, c = (t.detach().requires_grad_(True) for t in (q, k, v))
        swa(a, b, c, 2, shift, H, W, amask if shift else None).backward(wg)
    def fb_torch():
        a, b, c = (t.detach().requires_grad_(True) for t in (q, k, v))
        O.split_window_attention(a, b, c, 2, shift, H, W).backward(wg)
    print(f"shift={shift}: fwd+bwd ours {timeit(fb, 10):.1f} us; eager torch ops on the GPU {timeit(fb_torch, 5):.1f} us")
with torch.no_grad():
    for shift in (False, True):
        ours = swa(q, k, v, 2, shift, H, W, amask if shift else None)
        ref = O.split_window_attention(q.double(), k.double(), v.double(), 2, shift, H, W)
        err = ((ours.double() - ref).norm() / ref.norm()).item()
        t = timeit(lambda: swa(q, k, v, 2, shift, H, W, amask if shift else None))
        tt = timeit(lambda: O.split_window_attention(q, k, v, 2, shift, H, W), iters=5)
        print(f"shift={shift}: ours {t:.1f} us (rel-L2 {err:.2e}); eager torch ops on the GPU {tt:.1f} us")
