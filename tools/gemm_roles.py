#!/usr/bin/env python
"""Role wait-cycle profile of gemm_tc_kernel on the token-row linear layers: gemm_roles.py [pairs]."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from emip_b200 import _lib
from emip_b200.transformer_layer import linear_tm, linear_ln_tm, mlp_tm, linear_tm_multi

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
L = _lib.lib()
rows = 2 * pairs * 1936
x = torch.randn(rows, 128, device=dev)
x2 = torch.randn(rows, 256, device=dev)
w = torch.randn(128, 128, device=dev) / 11
w1 = torch.randn(1024, 256, device=dev) / 16
w2 = torch.randn(128, 1024, device=dev) / 32
g, b = torch.ones(128, device=dev), torch.zeros(128, device=dev)
h = torch.randn(rows, 1024, device=dev)
prof = torch.zeros(148, 8, dtype=torch.int64, device=dev)
cases = {
    "linear 128->128 (plain fp32 epilogue)": lambda: linear_tm(x, w),
    "linear 128->128 + LN + residual": lambda: linear_ln_tm(x, w, g, b, 1e-5, residual=x),
    "linear 1024->128 + LN + residual (= the mlp[2] launch)": lambda: linear_ln_tm(h, w2, g, b, 1e-5, residual=x),
    "mlp 256->1024->128 + LN (SUM of the mlp[0] + GELU / split launch and the mlp[2] launch)": lambda: mlp_tm(x2, w1, w2, g, b, 1e-5, residual=x),
}
with torch.no_grad():
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        L.emip_gemm_tc_set_profile_buffer(ctypes.c_void_p(prof.data_ptr()))
        prof.zero_()
        fn()
        torch.cuda.synchronize()
        L.emip_gemm_tc_set_profile_buffer(ctypes.c_void_p(0))
        p = prof.double().mean(0).tolist()
        print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call | producer total {p[0]:.0f} cyc, wait empty {p[1]:.0f} | "
              f"issuer total {p[2]:.0f}, wait full {p[3]:.0f}, wait acc_empty {p[4]:.0f} | epilogue warp 0 total {p[5]:.0f}, wait acc_full {p[6]:.0f}, "
              f"tiles {p[7]:.1f} -> {(p[5] - p[6]) / max(p[7], 1):.0f} cyc of epilogue work per tile")
