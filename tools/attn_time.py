"""Accuracy + timing of the fused attention kernel (attn_tc.cu) on the f2 and a5 shapes, incl. the lazy-rescale path."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
import cases
from emip_b200.window_attn import attention
from emip_b200.memory import Memory


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def ref_attn(q, k, v):
    q, k, v = q.double(), k.double(), v.double()
    return torch.softmax(q @ k.transpose(1, 2) / q.shape[-1] ** 0.5, -1) @ v


def timeit(fn, iters=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


g = torch.Generator().manual_seed(5)
for nb, n, scale, name in ((128, 484, 1.0, "f2 window"), (32, 1936, 1.0, "f2 full"), (64, 242, 1.0, "f2 22x11"), (3, 200, 4.0, "peaky"),
                           (2, 700, 1.0, "ramp")):
    q = torch.randn(nb, n, 128, generator=g) * scale
    k = torch.randn(nb, n, 128, generator=g) * scale
    v = torch.randn(nb, n, 128, generator=g)
    if name == "ramp":
        # row maxima that keep growing by > 2^16 from key tile to key tile: forces the O rescale path every tile
        k = k * 0.05 + torch.arange(n).view(1, n, 1) * 0.02 * torch.sign(q[:, :1, :])
        q = q.abs() * torch.sign(q[:, :1, :])
    q, k, v = q.cuda(), k.cuda(), v.cuda()
    with torch.no_grad():
        out = attention(q, k, v)
        r = ref_attn(q, k, v)
        s = (q.double() @ k.double().transpose(1, 2) / 128 ** 0.5)
        t = timeit(lambda: attention(q, k, v))
    print(f"{name:10s} nb={nb:4d} n={n:5d}: rel-L2 {rel(out, r):.2e}  score range [{s.min().item():.0f}, {s.max().item():.0f}]  {t:.1f} us", flush=True)

for T in (1, 3, 5):
    d = cases.a5_inputs(dict(b=1, t=T, h=44, w=44, scale=1.5, seed=57))
    t = {k_: d[k_].cuda() for k_ in ("m_in", "m_out", "q_in", "q_out")}
    outs = {}
    for exact in (True, False):
        m = Memory(); m.exact_fp32 = exact
        with torch.no_grad():
            outs[exact] = m(t["m_in"], t["m_out"], t["q_in"], t["q_out"])
            us = timeit(lambda: m(t["m_in"], t["m_out"], t["q_in"], t["q_out"]))
        print(f"a5 T={T} exact={exact}: {us:.1f} us", flush=True)
    o0, o1 = outs[True], outs[False]
    o0 = o0[0] if isinstance(o0, (tuple, list)) else o0
    o1 = o1[0] if isinstance(o1, (tuple, list)) else o1
    print(f"   tc vs exact rel-L2 {rel(o1, o0):.2e}")
