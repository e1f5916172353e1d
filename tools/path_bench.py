#!/usr/bin/env python
"""Per-row benchmark of the hot path (SURVEY.md section 8a rows a1..a5 and the 8f rows built so far), forward and
backward, on one B200.

For every row: CUDA-event time of the public op (inputs resident in HBM, L2 defeated by rotating input sets where the
footprint is small), algorithmic FLOPs / bytes, the fraction of the measured peak that bounds it, and the CPU oracle
(the reference algorithm restated, torch CPU fp32, all host threads) timed on a bounded sample beside it.

    python tools/path_bench.py [--no-cpu] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402
import cases  # noqa: E402

N, C, H, W = 1936, 128, 44, 44


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def gpu_time(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cpu_time(fn, budget_s=6.0):
    fn()
    t0, n = time.perf_counter(), 0
    while True:
        fn()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 5:
            break
    return (time.perf_counter() - t0) / n * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from oracle import restate as O
    from emip_b200.matching import global_correlation_softmax
    from emip_b200.flow_attn import FeatureFlowAttention
    from emip_b200.warp import flow_warp
    from emip_b200.injector import Injector
    from emip_b200.memory import Memory
    hbm, tf, src = peaks()
    dev = torch.device("cuda", 0)
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator(device=dev).manual_seed(3)
    rows = []

    def add(name, unit_n, unit, ms_f, ms_b, flops_f, flops_b, bytes_f, bytes_b, bound, cpu_f=None, cpu_b=None, note=""):
        r = {"row": name, "units": unit_n, "unit": unit, "fwd_ms": ms_f, "bwd_ms": ms_b,
             "fwd_units_per_s": unit_n / (ms_f * 1e-3), "bound": bound, "note": note}
        if bound == "tensor":
            r["fwd_frac_of_peak"] = flops_f / (ms_f * 1e-3) / 1e12 / tf
            r["fwd_alg_tflops"] = flops_f / (ms_f * 1e-3) / 1e12
        else:
            r["fwd_frac_of_peak"] = bytes_f / (ms_f * 1e-3) / 1e9 / hbm
            r["fwd_alg_gbs"] = bytes_f / (ms_f * 1e-3) / 1e9
        if ms_b is not None:
            r["bwd_alg_tflops"] = flops_b / (ms_b * 1e-3) / 1e12
            r["bwd_alg_gbs"] = bytes_b / (ms_b * 1e-3) / 1e9
        if cpu_f is not None:
            r["cpu_fwd_ms"] = cpu_f
            r["cpu_bwd_ms"] = cpu_b
            r["speedup_fwd"] = cpu_f / ms_f
            if cpu_b and ms_b:
                r["speedup_fwd_bwd"] = (cpu_f + cpu_b) / (ms_f + ms_b)
        rows.append(r)
        print(json.dumps(r))

    # ---------------- a1: global matching, B = 16 (config c2) ----------------
    B = 16
    sets = [(4.1 * torch.randn(B, C, H, W, device=dev, generator=g), 4.1 * torch.randn(B, C, H, W, device=dev, generator=g))
            for _ in range(4)]
    wf = torch.randn(2 * B, 2, H, W, device=dev, generator=g)
    wc = 0.05 * torch.randn(B, N, H, W, device=dev, generator=g)
    it = [0]

    def a1_fwd():
        f0, f1 = sets[it[0] % 4]
        it[0] += 1
        with torch.no_grad():
            return global_correlation_softmax(f0, f1, True)

    def a1_fwd_bwd():
        f0, f1 = sets[it[0] % 4]
        it[0] += 1
        a, b = f0.detach().requires_grad_(True), f1.detach().requires_grad_(True)
        flow, _, corr = global_correlation_softmax(a, b, True)
        torch.autograd.backward([flow, corr], [wf, wc])

    ms_f = gpu_time(a1_fwd)
    ms_fb = gpu_time(a1_fwd_bwd, iters=5)
    cf = cb = None
    if not args.no_cpu:
        c0, c1 = sets[0][0].cpu(), sets[0][1].cpu()
        cw, cc = wf.cpu(), wc.cpu()

        def c_f():
            with torch.no_grad():
                O.global_correlation_softmax(c0, c1, True)

        def c_fb():
            a, b = c0.clone().requires_grad_(True), c1.clone().requires_grad_(True)
            flow, _, corr = O.global_correlation_softmax(a, b, True)
            torch.autograd.backward([flow, corr], [cw, cc])
        cf = cpu_time(c_f)
        cb = cpu_time(c_fb) - cf
    fl = B * (2.0 * N * N * C + 8.0 * N * N)
    add("a1 global matching (bidir, corr emitted), B=16", B, "pairs", ms_f, ms_fb - ms_f, fl, 2 * fl + 2 * B * 2.0 * N * N * C * 2,
        B * (2 * C * N * 4 + 4 * N * 2 * 4 + N * N * 4), B * (3 * N * N * 4), "tensor", cf, cb,
        "bwd = exact-fp32 CUDA-core recompute (dflow and dcorr)")

    # ---------------- a2: flow-propagation attention, 2B = 32 ----------------
    B2 = 32
    m = FeatureFlowAttention(C).to(dev)
    xs = [4.1 * torch.randn(B2, C, H, W, device=dev, generator=g) for _ in range(4)]
    fl2 = 12.0 * torch.randn(B2, 2, H, W, device=dev, generator=g)
    wo = torch.randn(B2, 2, H, W, device=dev, generator=g)

    def a2_fwd():
        it[0] += 1
        with torch.no_grad():
            return m(xs[it[0] % 4], fl2)

    def a2_fwd_bwd():
        it[0] += 1
        x = xs[it[0] % 4].detach().requires_grad_(True)
        m(x, fl2).backward(wo)
    ms_f = gpu_time(a2_fwd)
    ms_fb = gpu_time(a2_fwd_bwd, iters=5)
    cf = cb = None
    if not args.no_cpu:
        prm = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        cx, cfl, cwo = xs[0].cpu(), fl2.cpu(), wo.cpu()

        def c_f():
            with torch.no_grad():
                O.feature_flow_attention(cx, cfl, prm["q_proj.weight"], prm["q_proj.bias"], prm["k_proj.weight"], prm["k_proj.bias"])

        def c_fb():
            x = cx.clone().requires_grad_(True)
            O.feature_flow_attention(x, cfl, prm["q_proj.weight"], prm["q_proj.bias"], prm["k_proj.weight"],
                                     prm["k_proj.bias"]).backward(cwo)
        cf = cpu_time(c_f)
        cb = cpu_time(c_fb) - cf
    fl = B2 * (2.0 * N * N * (C + 2) + 2 * 2.0 * N * C * C)
    add("a2 flow-propagation attention (incl. q/k Linear), 2B=32", B2 // 2, "pairs", ms_f, ms_fb - ms_f, fl, 2.5 * fl,
        B2 * (C * N * 4 + 4 * N * 4), B2 * (2 * C * N * 4), "tensor", cf, cb)

    # ---------------- a3: flow_warp, B = 64, 3 x 352 x 352 ----------------
    Bw, Cw, Hw, Ww = 64, 3, 352, 352
    x = torch.randn(Bw, Cw, Hw, Ww, device=dev, generator=g)
    base = 10.0 * torch.randn(Bw, 2, 1, 1, device=dev, generator=g)
    fl4 = torch.cat([cases.smooth_flow(11, Bw, Hw, Ww, 0.5), cases.smooth_flow(12, Bw, Hw, Ww, 0.5)], 1).to(dev)
    fl4[:, :2] += base
    fl4[:, 2:] -= base
    wout = torch.randn(Bw, Cw, Hw, Ww, device=dev, generator=g)

    def a3_fwd():
        with torch.no_grad():
            return flow_warp(x, fl4[:, 2:])

    def a3_fwd_bwd():
        f = fl4.detach().requires_grad_(True)
        flow_warp(x, f[:, 2:]).backward(wout)
    ms_f = gpu_time(a3_fwd)
    ms_fb = gpu_time(a3_fwd_bwd, iters=5)
    cf = cb = None
    if not args.no_cpu:
        sx, sf, sw = x[:8].cpu(), fl4[:8].cpu(), wout[:8].cpu()          # bounded sample: 8 of 64 images

        def c_f():
            with torch.no_grad():
                O.flow_warp(sx, sf[:, 2:])

        def c_fb():
            f = sf.clone().requires_grad_(True)
            O.flow_warp(sx, f[:, 2:]).backward(sw)
        cf = cpu_time(c_f) * 8
        cb = cpu_time(c_fb) * 8 - cf
    add("a3 flow_warp (model-like smooth flow), B=64 3x352x352", Bw, "images", ms_f, ms_fb - ms_f, 0, 0,
        Bw * Hw * Ww * (2 * Cw + 2) * 4, Bw * Hw * Ww * (2 * Cw + 4) * 4, "hbm", cf, cb,
        "bwd through autograd: includes the zero-fill + slice scatter of the [B,4,H,W] flow gradient; CPU sample = 8 images x 8")

    # ---------------- a4: injector, B = 16 ----------------
    B4 = 16
    inj = Injector().to(dev)
    inj.transformer.load_state_dict({k: v for k, v in cases.injector_params(7).items()})
    x4 = [2.2 * torch.randn(B4, C, H, W, device=dev, generator=g) for _ in range(4)]
    y4 = [torch.randn(B4, C, H, W, device=dev, generator=g) for _ in range(4)]
    w4 = torch.randn(B4, C, H, W, device=dev, generator=g)

    def a4_fwd():
        it[0] += 1
        with torch.no_grad():
            return inj(x4[it[0] % 4], y4[it[0] % 4])

    def a4_fwd_bwd():
        it[0] += 1
        a, b = x4[it[0] % 4].detach().requires_grad_(True), y4[it[0] % 4].detach().requires_grad_(True)
        inj(a, b).backward(w4)
        inj.zero_grad(set_to_none=True)
    ms_f = gpu_time(a4_fwd, iters=10)
    ms_fb = gpu_time(a4_fwd_bwd, iters=5)
    cf = cb = None
    if not args.no_cpu:
        prm = cases.injector_params(7)
        sx, sy, sw = x4[0][:4].cpu(), y4[0][:4].cpu(), w4[:4].cpu()    # bounded sample: 4 of 16

        def c_f():
            with torch.no_grad():
                O.injector(sx, sy, prm)

        def c_fb():
            a, b = sx.clone().requires_grad_(True), sy.clone().requires_grad_(True)
            pp = {k: v.clone().requires_grad_(True) for k, v in prm.items()}
            O.injector(a, b, pp).backward(sw)
        cf = cpu_time(c_f) * 4
        cb = cpu_time(c_fb) * 4 - cf
    fl = B4 * 0.86e9
    add("a4 prompt fusion (Injector), B=16", B4, "calls", ms_f, ms_fb - ms_f, fl, 2 * fl, B4 * 3 * C * N * 4, B4 * 5 * C * N * 4,
        "hbm", cf, cb, "1x1-conv GEMMs on tcgen05 (split-bf16, fp32 accumulate), the rest exact fp32 CUDA cores; 0.86 GFLOP and 2.97 MB algorithmic per sample; CPU sample = 4 x 4")

    # ---------------- a5: memory read, B = 1, T = 5 ----------------
    d5 = cases.a5_inputs(dict(b=1, t=5, h=44, w=44, scale=1.5, seed=57))
    t5 = {k: d5[k].to(dev) for k in ("m_in", "m_out", "q_in", "q_out")}
    w5 = d5["wout"].to(dev)
    mem = Memory()

    def a5_fwd():
        with torch.no_grad():
            return mem(t5["m_in"], t5["m_out"], t5["q_in"], t5["q_out"])

    def a5_fwd_bwd():
        tt = {k: v.detach().requires_grad_(True) for k, v in t5.items()}
        mem(tt["m_in"], tt["m_out"], tt["q_in"], tt["q_out"])[0].backward(w5)
    ms_f = gpu_time(a5_fwd, iters=10)
    ms_fb = gpu_time(a5_fwd_bwd, iters=5)
    cf = cb = None
    if not args.no_cpu:
        def c_f():
            with torch.no_grad():
                O.memory_read(d5["m_in"], d5["m_out"], d5["q_in"], d5["q_out"])

        def c_fb():
            tt = {k: d5[k].clone().requires_grad_(True) for k in ("m_in", "m_out", "q_in", "q_out")}
            O.memory_read(tt["m_in"], tt["m_out"], tt["q_in"], tt["q_out"])[0].backward(d5["wout"])
        cf = cpu_time(c_f)
        cb = cpu_time(c_fb) - cf
    M = 5 * N
    fl = 2.0 * M * N * 256
    add("a5 EMIP_long memory read, B=1 T=5 (9680 slots)", 1, "frames", ms_f, ms_fb - ms_f, fl, 2.5 * fl, (2 * M + 3 * N) * 128 * 4,
        (4 * M + 4 * N) * 128 * 4, "tensor", cf, cb, "forward: one fused flash-style tcgen05 kernel; backward: three launches of the tcgen05 gradient kernel (dQ; dK; dV); split-bf16, fp32 accumulate")

    # ---------------- f1: conv_corr[0] on the never-materialised cost volume, B = 16 ----------------
    from emip_b200.conv_corr import conv_corr_first_layer
    Of = 968
    wq = torch.randn(Of, N, 3, 3, device=dev, generator=g) * (9 * N) ** -0.5
    bq = 0.1 * torch.randn(Of, device=dev, generator=g)

    def f1_fwd():
        f0, f1 = sets[it[0] % 4]
        it[0] += 1
        with torch.no_grad():
            return conv_corr_first_layer(f0, f1, wq, bq)
    ms_f = gpu_time(f1_fwd, iters=10)
    wqg, bqg = wq.clone().requires_grad_(True), bq.clone().requires_grad_(True)
    wof = torch.randn(16, Of, H, W, device=dev, generator=g)

    def f1_fwd_bwd():
        f0, f1 = sets[it[0] % 4]
        it[0] += 1
        a, b = f0.detach().requires_grad_(True), f1.detach().requires_grad_(True)
        conv_corr_first_layer(a, b, wqg, bqg).backward(wof)
        wqg.grad = None
        bqg.grad = None
    ms_f1fb = gpu_time(f1_fwd_bwd, iters=5)
    cf = None
    if not args.no_cpu:
        c0, c1, cwq, cbq = sets[0][0][:1].cpu(), sets[0][1][:1].cpu(), wq.cpu(), bq.cpu()     # bounded sample: 1 of 16

        def c_f():
            with torch.no_grad():
                O.conv_corr_first_layer(c0, c1, cwq, cbq)
        cf = cpu_time(c_f) * 16
    fl = 16 * 2.0 * (Of * 9 * N * C + Of * N * 9 * C)
    add("f1 conv_corr[0] on the never-materialised cost volume, B=16 (968 out channels)", 16, "pairs", ms_f, ms_f1fb - ms_f, fl, 2.5 * fl,
        16 * (2 * C * N * 4 + Of * N * 4), 0, "tensor", cf, None,
        "two split-bf16 tcgen05 GEMMs (8.6 GFLOP/sample) instead of a 65.3 GFLOP/sample convolution over corr; "
        "CPU = the reference composition (matmul + conv2d on corr), sample = 1 pair x 16; backward = five split-bf16 tcgen05 GEMMs (tools/f1_bwd_time.py)")

    # ---------------- f3: occlusion mask, f4: convex upsampling ----------------
    from emip_b200.warp import get_occu_mask_backward
    from emip_b200.upsample import upsample_flow_convex

    def f3_fwd():
        return get_occu_mask_backward(fl4[:, 2:])
    ms_f = gpu_time(f3_fwd)
    cf = None
    if not args.no_cpu:
        sf = fl4[:8, 2:].cpu()
        cf = cpu_time(lambda: O.occu_mask_backward(sf, 0.2)) * 8
    add("f3 backward-flow occlusion mask, B=64 352x352", Bw, "images", ms_f, None, 0, 0, Bw * Hw * Ww * 3 * 4, 0, "hbm", cf, None,
        "bilinear splat (4 atomicAdd per pixel) + threshold; CPU sample = 8 images x 8")
    Bu = 32
    cfl = 12.0 * torch.randn(Bu, 2, H, W, device=dev, generator=g)
    cmask = torch.randn(Bu, 576, H, W, device=dev, generator=g)
    wup = torch.randn(Bu, 2, 8 * H, 8 * W, device=dev, generator=g)

    def f4_fwd():
        with torch.no_grad():
            return upsample_flow_convex(cfl, cmask)

    def f4_fwd_bwd():
        a, b = cfl.detach().requires_grad_(True), cmask.detach().requires_grad_(True)
        upsample_flow_convex(a, b).backward(wup)
    ms_f = gpu_time(f4_fwd)
    ms_fb = gpu_time(f4_fwd_bwd, iters=5)
    cf = cb = None
    if not args.no_cpu:
        sfl, smk, swu = cfl[:4].cpu(), cmask[:4].cpu(), wup[:4].cpu()

        def c_f():
            with torch.no_grad():
                O.upsample_flow_convex(sfl, smk)

        def c_fb():
            a, b = sfl.clone().requires_grad_(True), smk.clone().requires_grad_(True)
            O.upsample_flow_convex(a, b).backward(swu)
        cf = cpu_time(c_f) * 8
        cb = cpu_time(c_fb) * 8 - cf
    add("f4 convex x8 flow upsampling, 2B=32", Bu, "flows", ms_f, ms_fb - ms_f, 0, 0, Bu * (576 * N + 2 * N + 2 * 64 * N) * 4,
        Bu * (2 * 576 * N + 2 * N + 2 * 64 * N) * 4, "hbm", cf, cb, "CPU sample = 4 flows x 8")

    # ---------------- f2: split-window attention of the FeatureTransformer, 2B = 32 ----------------
    from emip_b200.window_attn import single_head_split_window_attention
    qw, kw, vw = (2.0 * torch.randn(32, N, C, device=dev, generator=g) for _ in range(3))
    amask = O.shift_window_attn_mask(H, W, 22, 22, 11, 11).to(dev)
    wf2 = torch.randn(32, N, C, device=dev, generator=g)
    res2 = {}
    for shift in (False, True):
        def f2_fwd():
            with torch.no_grad():
                return single_head_split_window_attention(qw, kw, vw, 2, shift, H, W, amask if shift else None)

        def f2_torch():                                   # the reference's op sequence on the same GPU (library kernels)
            with torch.no_grad():
                return O.split_window_attention(qw, kw, vw, 2, shift, H, W)
        def f2_fwd_bwd():
            a, b, c = (t.detach().requires_grad_(True) for t in (qw, kw, vw))
            single_head_split_window_attention(a, b, c, 2, shift, H, W, amask if shift else None).backward(wf2)
        res2[shift] = (gpu_time(f2_fwd, iters=10), gpu_time(f2_torch, iters=5, warm=2), gpu_time(f2_fwd_bwd, iters=5))
    cf = None
    if not args.no_cpu:
        cq, ck, cv = qw[:4].cpu(), kw[:4].cpu(), vw[:4].cpu()
        cf = cpu_time(lambda: O.split_window_attention(cq, ck, cv, 2, True, H, W)) * 8
    fl = 32 * 4 * 2.0 * 484 * 484 * (C + C)
    add("f2 split-window attention (one shifted layer), 2B=32", 16, "pairs", res2[True][0], res2[True][2] - res2[True][0], fl, 2.5 * fl,
        32 * 4 * N * C * 4, 0, "tensor", cf, None,
        f"one C-ABI call per layer (fused flash kernel, window gather / scatter inside); unshifted layer {res2[False][0]:.3f} ms fwd, "
        f"{res2[False][2] - res2[False][0]:.3f} ms bwd; backward = tcgen05 gradient kernel per block group (+ torch gather / scatter "
        f"copies); the same formula as eager torch ops on this GPU: shifted {res2[True][1]:.3f} ms, unshifted {res2[False][1]:.3f} ms "
        f"forward; 12 such layers per pair of feature maps; CPU sample = 4 x 8")

    # ---------------- f2b: one whole FeatureTransformer block pair (self layer + cross / FFN layer), 2B = 32 ----------------
    from types import SimpleNamespace as NS
    from emip_b200.transformer_layer import transformer_layer_forward
    pb = {k: v.to(dev) for k, v in cases.f2b_inputs(cases.F2B_CASES["f2b_full"])["params"].items()}
    lin = lambda n: NS(weight=pb[n + ".weight"])
    lnm = lambda n: NS(weight=pb[n + ".weight"], bias=pb[n + ".bias"], eps=1e-5)
    mk = lambda no_ffn: NS(attention_type="swin", nhead=1, no_ffn=no_ffn, with_shift=True, q_proj=lin("q_proj"), k_proj=lin("k_proj"),
                           v_proj=lin("v_proj"), merge=lin("merge"), norm1=lnm("norm1"), norm2=lnm("norm2"),
                           mlp=[lin("mlp.0"), None, lin("mlp.2")])
    l_self, l_cross = mk(True), mk(False)
    src_b, tgt_b = torch.randn(32, N, C, device=dev, generator=g), torch.randn(32, N, C, device=dev, generator=g)

    def block(s_, t_):
        s_ = transformer_layer_forward(l_self, s_, s_, height=H, width=W, shifted_window_attn_mask=amask, attn_num_splits=2)
        return transformer_layer_forward(l_cross, s_, t_, height=H, width=W, shifted_window_attn_mask=amask, attn_num_splits=2)

    def f2b_fwd():
        with torch.no_grad():
            return block(src_b, tgt_b)

    def f2b_fwd_bwd():
        a, b = src_b.detach().requires_grad_(True), tgt_b.detach().requires_grad_(True)
        block(a, b).backward(wf2)

    def f2b_torch():                                      # the reference's op sequence on the same GPU (library kernels)
        with torch.no_grad():
            s_ = O.transformer_layer(src_b, src_b, pb, True, 2, True, H, W)
            return O.transformer_layer(s_, tgt_b, pb, False, 2, True, H, W)
    tb_f, tb_fb, tb_t = gpu_time(f2b_fwd, iters=10), gpu_time(f2b_fwd_bwd, iters=5), gpu_time(f2b_torch, iters=3, warm=1)
    cf = None
    if not args.no_cpu:
        cs_, ct_ = src_b[:2].cpu(), tgt_b[:2].cpu()
        pc = {k: v.cpu() for k, v in pb.items()}
        cf = cpu_time(lambda: O.transformer_layer(O.transformer_layer(cs_, cs_, pc, True, 2, True, H, W), ct_, pc, False, 2, True, H, W)) * 16
    ntok = 32 * N
    fl = 2.0 * ntok * (8 * C * C + 2 * C * 8 * C + 8 * C * C) + 2 * 32 * 4 * 2.0 * 484 * 484 * (C + C)
    add("f2b FeatureTransformer block pair (self + cross / FFN layer, shifted), 2B=32", 16, "pairs", tb_f, tb_fb - tb_f, fl, 2.0 * fl,
        ntok * C * 4 * 3, 0, "tensor", cf, None,
        f"TransformerLayer.forward as the drop-in: ten split-bf16 tcgen05 linear layers (executed MMA FLOPs = 3 x algorithmic), "
        f"fused MLP call (GELU + operand split in the first GEMM's epilogue), two window-attention calls, three LayerNorm + residual "
        f"launches; the reference's op sequence in eager torch on this GPU (cuBLAS fp32, library attention): {tb_t:.3f} ms forward; "
        f"six block pairs per forward; CPU sample = 2 maps x 16")

    # ---------------- f3b: photometric loss term (L1 + SSIM 3x3, masked), B = 64 ----------------
    from emip_b200.photometric import photometric_loss
    rec0 = x + 0.2 * torch.randn(Bw, Cw, Hw, Ww, device=dev, generator=g)
    occ = (torch.rand(Bw, 1, Hw, Ww, device=dev, generator=g) > 0.2).float()

    def f3b_fwd():
        with torch.no_grad():
            return photometric_loss(x, rec0, occ)

    def f3b_fwd_bwd():
        r = rec0.detach().requires_grad_(True)
        photometric_loss(x, r, occ).backward()

    def torch_fwd_bwd():                                  # the reference's op sequence on the same GPU (library kernels)
        r = rec0.detach().requires_grad_(True)
        O.photometric_loss(x, r, occ).backward()
    ms_f = gpu_time(f3b_fwd)
    ms_fb = gpu_time(f3b_fwd_bwd, iters=5)
    ms_torch = gpu_time(torch_fwd_bwd, iters=3, warm=1)
    cf = cb = None
    if not args.no_cpu:
        sx, sr, so = x[:8].cpu(), rec0[:8].cpu(), occ[:8].cpu()

        def c_f():
            with torch.no_grad():
                O.photometric_loss(sx, sr, so)

        def c_fb():
            r = sr.clone().requires_grad_(True)
            O.photometric_loss(sx, r, so).backward()
        cf = cpu_time(c_f) * 8
        cb = cpu_time(c_fb) * 8 - cf
    add("f3b photometric loss term (L1 + SSIM), B=64 3x352x352", Bw, "images", ms_f, ms_fb - ms_f, 0, 0,
        Bw * Hw * Ww * (2 * Cw + 1) * 4, Bw * Hw * Ww * (3 * Cw + 1) * 4, "hbm", cf, cb,
        f"the same formula as eager torch ops on this GPU (what the reference launches): fwd+bwd {ms_torch:.2f} ms; CPU sample = 8 x 8")

    out = {"peaks": {"hbm_gbs": hbm, "bf16_tflops": tf, "source": src}, "cpu_threads": os.cpu_count(), "rows": rows}
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
