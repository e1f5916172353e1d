"""One FeatureTransformer block pair (self-attention layer + cross-attention / FFN layer, transformer.py:108-196) at
2B = 32 feature maps of 44x44 tokens: emip_b200.transformer_layer against the reference formulation in eager torch on
the same GPU.  Prints per-layer times (forward, forward + backward w.r.t. the token rows) and per-call linear times."""
import json
import sys

import torch

sys.path.insert(0, ".")
from emip_b200.transformer_layer import transformer_layer_forward, linear_tm, layer_norm_tm, mlp_tm   # noqa: E402
from emip_b200 import window_attn                                                                # noqa: E402

B2, H, W, C = 32, 44, 44, 128
nn = torch.nn


class Layer(nn.Module):
    def __init__(self, no_ffn, with_shift):
        super().__init__()
        self.attention_type, self.nhead, self.no_ffn, self.with_shift = "swin", 1, no_ffn, with_shift
        self.q_proj, self.k_proj, self.v_proj, self.merge = (nn.Linear(C, C, bias=False) for _ in range(4))
        self.norm1 = nn.LayerNorm(C)
        if not no_ffn:
            self.mlp = nn.Sequential(nn.Linear(2 * C, 8 * C, bias=False), nn.GELU(), nn.Linear(8 * C, C, bias=False))
            self.norm2 = nn.LayerNorm(C)

    def eager(self, source, target):                     # the reference's op sequence, attention core = ours (measured in f2)
        q, k, v = self.q_proj(source), self.k_proj(target), self.v_proj(target)
        m = window_attn.single_head_split_window_attention(q, k, v, 2, self.with_shift, H, W, torch.zeros(1, device="cuda"))
        m = self.norm1(self.merge(m))
        if not self.no_ffn:
            m = self.norm2(self.mlp(torch.cat([source, m], -1)))
        return source + m

    def ours(self, source, target):
        return transformer_layer_forward(self, source, target, height=H, width=W,
                                         shifted_window_attn_mask=torch.zeros(1, device="cuda"), attn_num_splits=2)


def t(fn, it=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3


torch.manual_seed(0)
res = {"allow_tf32": torch.backends.cuda.matmul.allow_tf32, "tokens": B2 * H * W}
src = torch.randn(B2, H * W, C, device="cuda")
tgt = torch.randn(B2, H * W, C, device="cuda")
wo = torch.randn(B2, H * W, C, device="cuda")
for no_ffn in (True, False):
    lay = Layer(no_ffn, False).cuda().requires_grad_(False)
    name = "self" if no_ffn else "cross_ffn"
    with torch.no_grad():
        res[name + "_fwd_us"] = {"ours": t(lambda: lay.ours(src, tgt)), "eager": t(lambda: lay.eager(src, tgt))}
        res[name + "_rel_l2"] = ((lay.ours(src, tgt) - lay.eager(src, tgt)).norm() / lay.eager(src, tgt).norm()).item()
    s, g = src.clone().requires_grad_(True), tgt.clone().requires_grad_(True)

    def fb(f):
        s.grad = None
        g.grad = None
        f(s, g).backward(wo)
    res[name + "_fwd_bwd_us"] = {"ours": t(lambda: fb(lay.ours)), "eager": t(lambda: fb(lay.eager))}
x2 = src.view(-1, C)
h = torch.randn(B2 * H * W, 8 * C, device="cuda")
xc = torch.cat([x2, x2], -1)
w1, w2, w3 = (torch.randn(o, k, device="cuda") * k ** -0.5 for o, k in ((C, C), (8 * C, 2 * C), (C, 8 * C)))
g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
with torch.no_grad():
    res["linear_128_128_us"] = {"ours": t(lambda: linear_tm(x2, w1)), "eager": t(lambda: x2 @ w1.T)}
    res["linear_256_1024_us"] = {"ours": t(lambda: linear_tm(xc, w2)), "eager": t(lambda: xc @ w2.T)}
    res["gelu_linear_1024_128_us"] = {"ours": t(lambda: linear_tm(h, w3, gelu_in=True)),
                                      "eager": t(lambda: torch.nn.functional.gelu(h) @ w3.T)}
    res["mlp_256_1024_128_fused_us"] = {"ours": t(lambda: mlp_tm(xc, w2, w3)),
                                        "eager": t(lambda: torch.nn.functional.gelu(xc @ w2.T) @ w3.T)}
    res["ln_residual_us"] = {"ours": t(lambda: layer_norm_tm(x2, g, b, 1e-5, residual=x2)),
                             "eager": t(lambda: x2 + torch.nn.functional.layer_norm(x2, (C,), g, b))}
print(json.dumps(res, indent=1))
