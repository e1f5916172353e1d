"""Accuracy + timing of the tensor-core attention backward (attn_bwd_tc.cu) on f2 / a5 shapes."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
import cases
from emip_b200.window_attn import attention
from emip_b200.memory import Memory


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def timeit(fn, iters=10):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


g = torch.Generator().manual_seed(7)
for nb, n, scale in ((2, 100, 1.0), (3, 300, 1.5), (128, 484, 1.0), (8, 1936, 1.0)):
    q, k, v, w = (torch.randn(nb, n, 128, generator=g) * (scale if i < 2 else 1.0) for i in range(4))
    q, k, v, w = q.cuda(), k.cuda(), v.cuda(), w.cuda()
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    ref = torch.softmax(qd @ kd.transpose(1, 2) / 128 ** 0.5, -1) @ vd
    ref.backward(w.double())
    qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = attention(qq, kk, vv)
    out.backward(w)
    torch.cuda.synchronize()
    def fb():
        a, b, c = (t.detach().requires_grad_(True) for t in (q, k, v))
        attention(a, b, c).backward(w)
    def f():
        with torch.no_grad(): attention(q, k, v)
    tf, tfb = timeit(f), timeit(fb)
    print(f"nb={nb:4d} n={n:5d}: out {rel(out, ref):.1e} dq {rel(qq.grad, qd.grad):.1e} dk {rel(kk.grad, kd.grad):.1e} dv {rel(vv.grad, vd.grad):.1e}"
          f"   fwd {tf:.0f} us, fwd+bwd {tfb:.0f} us", flush=True)

from torch.profiler import profile, ProfilerActivity
for T in (1, 5):
    d = cases.a5_inputs(dict(b=1, t=T, h=44, w=44, scale=1.5, seed=57))
    dv_ = {k_: d[k_].cuda() for k_ in ("m_in", "m_out", "q_in", "q_out", "wout")}
    grads = {}
    for exact in (True, False):
        t = {k_: dv_[k_].clone().requires_grad_(True) for k_ in ("m_in", "m_out", "q_in", "q_out")}
        m = Memory(); m.exact_fp32 = exact
        out, _ = m(t["m_in"], t["m_out"], t["q_in"], t["q_out"])
        (out * dv_["wout"]).sum().backward()
        grads[exact] = {k_: v_.grad.clone() for k_, v_ in t.items()}
        tt = {k_: dv_[k_].clone().requires_grad_(True) for k_ in ("m_in", "m_out", "q_in", "q_out")}
        def f():
            with torch.no_grad(): m(dv_["m_in"], dv_["m_out"], dv_["q_in"], dv_["q_out"])
        def fb():
            o, _ = m(tt["m_in"], tt["m_out"], tt["q_in"], tt["q_out"])
            o.backward(dv_["wout"])
        print(f"a5 T={T} exact={exact}: fwd {timeit(f):.0f} us, fwd+bwd {timeit(fb):.0f} us", flush=True)
        if T == 5 and not exact:
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(3): fb()
                torch.cuda.synchronize()
            for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:10]:
                print(f"      {e.key[:70]:70s} n={e.count // 3:3d}/call {e.device_time_total / 3:8.1f} us/call")
    print("   tc vs exact:", {k_: f"{rel(grads[False][k_], grads[True][k_]):.1e}" for k_ in grads[True]})
q, k, v, w = (torch.randn(128, 484, 128, generator=g).cuda() for _ in range(4))
def fb2():
    a, b, c = (t.detach().requires_grad_(True) for t in (q, k, v))
    attention(a, b, c).backward(w)
fb2()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): fb2()
    torch.cuda.synchronize()
print("f2 window shape fwd+bwd:")
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:10]:
    print(f"      {e.key[:70]:70s} n={e.count // 3:3d}/call {e.device_time_total / 3:8.1f} us/call")
