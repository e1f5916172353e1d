import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
import cases
from emip_b200.memory import Memory
for T in (1, 3, 5):
    d = cases.a5_inputs(dict(b=1, t=T, h=44, w=44, scale=1.5, seed=57))
    t = {k: d[k].cuda() for k in ("m_in", "m_out", "q_in", "q_out")}
    for exact in (True, False):
        m = Memory(); m.exact_fp32 = exact
        with torch.no_grad():
            for _ in range(3): m(t["m_in"], t["m_out"], t["q_in"], t["q_out"])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): m(t["m_in"], t["m_out"], t["q_in"], t["q_out"])
            e1.record(); torch.cuda.synchronize()
        print(f"T={T} exact={exact}: {e0.elapsed_time(e1)/20*1e3:.1f} us")
