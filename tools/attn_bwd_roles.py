"""Per-role cycle breakdown of the attention backward kernel, one launch mode at a time (diagnostic; GPU box)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from emip_b200 import _lib
from emip_b200.window_attn import _attention_bwd
L = _lib.lib()
names = ["math:wait s_full", "math:tab+bars", "math:tmem ld", "math:W math", "math:wait w_empty", "math:st+arrive", "math:wait acc_full",
         "math:total", "mma:wait s_empty", "mma:wait ring", "mma:wait w_full", "mma:wait acc_empty", "mma:wait x_full", "mma:total"]
nb, n = 128, 484
g = torch.Generator().manual_seed(5)
q, k, v, w = (torch.randn(nb, n, 128, generator=g).cuda() for _ in range(4))
prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
_attention_bwd(q, k, v, w)
torch.cuda.synchronize()
L.emip_attn_tc_set_profile_buffer(ctypes.c_void_p(prof.data_ptr()))
_attention_bwd(q, k, v, w)      # the buffer keeps the LAST launch: PV (dV); fwd and ROW / COL overwrite earlier
torch.cuda.synchronize()
L.emip_attn_tc_set_profile_buffer(ctypes.c_void_p(0))
p = prof.view(148, 16).double().cpu()
items = nb * 4
full = p[: items % 148 or 148].mean(0)
tiles = -(-items // 148) * 4
print(f"== PV launch, nb={nb} n={n}: CTAs with {-(-items // 148)} items, {tiles} tiles (mean cycles)")
for i, nm in enumerate(names):
    print(f"   {nm:20s} {full[i]:10.0f}   per tile {full[i] / tiles:8.0f}")
