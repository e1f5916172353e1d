"""Per-role cycle breakdown of the fused attention kernel (diagnostic; run on the GPU box)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from emip_b200 import _lib
from emip_b200.window_attn import attention
L = _lib.lib()
names = ["smx:wait s_full", "smx:tmem ld", "smx:max+xchg", "smx:exp+split", "smx:wait p_empty", "smx:st+arrive", "smx:wait o_full",
         "smx:total", "mma:wait s_empty", "mma:wait ring", "mma:wait p_full", "mma:wait o_empty", "mma:wait q_full", "mma:total"]
for nb, n in ((128, 484), (32, 1936)):
    g = torch.Generator().manual_seed(5)
    q, k, v = (torch.randn(nb, n, 128, generator=g).cuda() for _ in range(3))
    prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        for rep in range(3):
            L.emip_attn_tc_set_profile_buffer(ctypes.c_void_p(prof.data_ptr() if rep == 2 else 0))
            attention(q, k, v)
        torch.cuda.synchronize()
        L.emip_attn_tc_set_profile_buffer(ctypes.c_void_p(0))
    p = prof.view(148, 16).double().cpu()
    nrt, nkt = (n + 127) // 128, (n + 127) // 128
    items = nb * nrt
    full = p[: items % 148 or 148].mean(0)
    print(f"== nb={nb} n={n}: {items} items x {nkt} key tiles; CTAs with {-(-items // 148)} items (mean cycles)")
    tiles = -(-items // 148) * nkt
    for i, nm in enumerate(names):
        print(f"   {nm:18s} {full[i]:10.0f}   per tile {full[i] / tiles:8.0f}")
