"""GPU parity: CUDA kernels (through the C ABI) vs golden vectors recorded from the
reference and vs the CPU oracle on identical seeded inputs.  Run with -m gpu on a B200."""
import pytest
import torch

import cases
from oracle import restate as O

pytestmark = pytest.mark.gpu

# north_star: outputs within 1e-3 relative (rel-L2 per tensor) in fp32
TOL_OUT = 1e-3
TOL_EXACT = 5e-5       # exact-fp32 CUDA-core path: re-association only
TOL_GRAD = 1e-3


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def dev(t):
    return t.cuda()


# ----------------------------------------------------------------------------- a3
@pytest.mark.parametrize("name", list(cases.A3_CASES))
def test_a3_flow_warp_golden(golden, name):
    from emip_b200.warp import flow_warp
    g = golden(name)
    s = cases.A3_CASES[name]
    d = cases.a3_inputs(s)
    x = dev(d["x"]).requires_grad_(True)
    fl = dev(d["flow"]).requires_grad_(True)
    out = flow_warp(x, fl, pad=s["pad"])
    cases.check_packed(out, g["out"], 1e-5, "out")
    (out * dev(d["wout"])).sum().backward()
    cases.check_packed(fl.grad, g["dflow"], 1e-4, "dflow")
    cases.check_packed(x.grad, g["dx"], 1e-5, "dx")


def test_a3_flow_warp_slices(golden):
    from emip_b200.warp import flow_warp
    g = golden("a3_slice")
    s = g["spec"]
    flow4 = dev(cases.randn(s["seed"], (s["b"], 4, s["h"], s["w"]), s["sigma"]))
    x = dev(cases.randn(s["seed"] + 1, (s["b"], 3, s["h"], s["w"])))
    cases.check_packed(flow_warp(x, flow4[:, :2]), g["out_fw"], 1e-5)
    cases.check_packed(flow_warp(x, flow4[:, 2:]), g["out_bw"], 1e-5)


@pytest.mark.parametrize("kind,mag", [("iid", 5.0), ("iid", 20.0), ("smooth", 30.0)])
def test_a3_flow_warp_fullsize_vs_oracle(kind, mag):
    """352x352 (config c5 shape, B reduced) vs the CPU oracle, incl. backward to flow."""
    from emip_b200.warp import flow_warp
    B, C, H, W = 2, 3, 352, 352
    x = cases.randn(61, (B, C, H, W))
    fl4 = cases.randn(62, (B, 4, H, W), mag) if kind == "iid" else torch.cat(
        [cases.smooth_flow(63, B, H, W, mag), cases.smooth_flow(64, B, H, W, mag)], 1)
    wout = cases.randn(65, (B, C, H, W))
    flc = fl4[:, 2:].clone().requires_grad_(True)
    ref = O.flow_warp(x, flc)
    (ref * wout).sum().backward()
    flg = dev(fl4).requires_grad_(True)
    out = flow_warp(dev(x), flg[:, 2:])
    (out * dev(wout)).sum().backward()
    assert rel(out, ref) < 1e-5
    assert rel(flg.grad[:, 2:], flc.grad) < 1e-4
    assert flg.grad[:, :2].abs().max().item() == 0.0


def test_a3_properties_full_batch():
    """Size-independent properties at the benchmark size (B=64, 3x352x352)."""
    from emip_b200.warp import flow_warp
    B, C, H, W = 64, 3, 352, 352
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    zero = torch.zeros(B, 2, H, W, device="cuda")
    assert (flow_warp(x, zero) - x).abs().max().item() < 2e-3          # identity up to the reference's fp32 grid round trip (|ix - j| <~ 1e-4 px)
    shift = zero.clone()
    shift[:, 0] = 3.0                                                   # integer shift: out[..., j] = x[..., min(j+3, W-1)]
    out = flow_warp(x, shift)
    assert (out[..., : W - 3] - x[..., 3:]).abs().max().item() < 2e-3
    assert (out[..., W - 3:] - x[..., W - 1:]).abs().max().item() < 2e-3
    fl = 20 * torch.randn(B, 2, H, W, device="cuda", generator=g)
    a = torch.randn(B, 1, 1, 1, device="cuda", generator=g)
    x2 = torch.randn(B, C, H, W, device="cuda", generator=g)
    lhs = flow_warp(a * x + x2, fl)
    rhs = a * flow_warp(x, fl) + flow_warp(x2, fl)                      # linear in the image
    assert rel(lhs, rhs) < 1e-5
    assert out.min() >= x.min() - 1e-4 and out.max() <= x.max() + 1e-4  # convex combination (border mode)


def test_a3_empty_batch():
    from emip_b200.warp import flow_warp
    x = torch.zeros(0, 3, 8, 8, device="cuda")
    assert flow_warp(x, torch.zeros(0, 2, 8, 8, device="cuda")).shape == (0, 3, 8, 8)


# ----------------------------------------------------------------------------- a1
@pytest.mark.parametrize("exact", [True, False], ids=["exact_fp32", "tcgen05"])
@pytest.mark.parametrize("name", list(cases.A1_CASES))
def test_a1_global_matching_golden(golden, name, exact):
    from emip_b200.matching import global_correlation_softmax
    g = golden(name)
    s = cases.A1_CASES[name]
    d = cases.a1_inputs(s)
    f0 = dev(d["f0"]).requires_grad_(True)
    f1 = dev(d["f1"]).requires_grad_(True)
    flow, prob, corr = global_correlation_softmax(f0, f1, True, exact_fp32=exact)
    assert prob is None
    tol = TOL_EXACT if exact else TOL_OUT
    e_flow = cases.check_packed(flow, g["flow"], tol, "flow")
    e_corr = cases.check_packed(corr, g["corr"], TOL_EXACT, "corr")
    print(f"{name} exact={exact}: flow rel-L2 {e_flow:.2e} corr rel-L2 {e_corr:.2e}")
    ((flow * dev(d["wflow"])).sum() + (corr * dev(d["wcorr"])).sum()).backward()
    cases.check_packed(f0.grad, g["df0"], TOL_GRAD, "df0")
    cases.check_packed(f1.grad, g["df1"], TOL_GRAD, "df1")


@pytest.mark.parametrize("exact", [True, False], ids=["exact_fp32", "tcgen05"])
def test_a1_insitu_c1(golden, exact):
    """Features captured inside a real CoUpdater forward (config c1): logits 67..212, near-one-hot softmax."""
    from emip_b200.matching import global_correlation_softmax
    g = golden("c1_insitu")
    flow, _, corr = global_correlation_softmax(dev(g["f0"]), dev(g["f1"]), True, exact_fp32=exact)
    e = cases.check_packed(flow, g["flow"], TOL_EXACT if exact else TOL_OUT, "flow")
    cases.check_packed(corr, g["corr"], TOL_EXACT, "corr")
    print(f"c1 in-situ exact={exact}: flow rel-L2 {e:.2e}")


@pytest.mark.parametrize("exact", [True, False], ids=["exact_fp32", "tcgen05"])
def test_a1_unidirectional_and_no_corr(exact):
    from emip_b200.matching import global_correlation_softmax
    s = cases.A1_CASES["a1_ragged"]
    d = cases.a1_inputs(s)
    ref_flow, _, ref_corr = O.global_correlation_softmax(d["f0"], d["f1"], False)
    flow, _, corr = global_correlation_softmax(dev(d["f0"]), dev(d["f1"]), False, exact_fp32=exact)
    assert flow.shape == ref_flow.shape and corr.shape == ref_corr.shape
    assert rel(flow, ref_flow) < (TOL_EXACT if exact else TOL_OUT)
    assert rel(corr, ref_corr) < TOL_EXACT
    flow2, _, corr2 = global_correlation_softmax(dev(d["f0"]), dev(d["f1"]), True, return_corr=False, exact_fp32=exact)
    assert corr2 is None and rel(flow2[: s["b"]], ref_flow) < (TOL_EXACT if exact else TOL_OUT)


def test_a1_c2_properties():
    """Config c2 (B=16, 44x44, C=128): tensor-core path vs exact path, and swap symmetry."""
    from emip_b200.matching import global_correlation_softmax
    f0 = dev(cases.randn(2, (16, 128, 44, 44), 4.1))
    f1 = dev(cases.randn(3, (16, 128, 44, 44), 4.1))
    flow_tc, _, corr_tc = global_correlation_softmax(f0, f1, True)
    flow_ex, _, corr_ex = global_correlation_softmax(f0, f1, True, exact_fp32=True)
    assert rel(flow_tc, flow_ex) < TOL_OUT
    assert rel(corr_tc, corr_ex) < TOL_EXACT
    # swapping the frames swaps the forward and backward flows and transposes corr
    flow_sw, _, corr_sw = global_correlation_softmax(f1, f0, True)
    assert rel(flow_sw[:16], flow_tc[16:]) < 1e-6 and rel(flow_sw[16:], flow_tc[:16]) < 1e-6
    assert rel(corr_sw.view(16, 1936, 1936).transpose(1, 2), corr_tc.view(16, 1936, 1936)) < 1e-6


TOL_BF16 = 2e-2        # north_star: bf16 mode tolerance


@pytest.mark.parametrize("name", ["a1_full", "a1_lowcontrast", "a1_ragged"])
def test_a1_bf16_mode(golden, name):
    """bf16 inference mode (single-pass bf16 operands): within 2e-2 of the fp32 reference."""
    from emip_b200.matching import global_correlation_softmax
    g = golden(name)
    d = cases.a1_inputs(cases.A1_CASES[name])
    flow, _, corr = global_correlation_softmax(dev(d["f0"]), dev(d["f1"]), True, bf16=True)
    e_flow = cases.check_packed(flow, g["flow"], TOL_BF16, "flow")
    e_corr = cases.check_packed(corr, g["corr"], TOL_BF16, "corr")
    print(f"{name} bf16: flow rel-L2 {e_flow:.2e} corr rel-L2 {e_corr:.2e}")


def test_a1_bf16_insitu_c1(golden):
    from emip_b200.matching import global_correlation_softmax
    g = golden("c1_insitu")
    flow, _, corr = global_correlation_softmax(dev(g["f0"]), dev(g["f1"]), True, bf16=True)
    e = cases.check_packed(flow, g["flow"], TOL_BF16, "flow")
    print(f"c1 in-situ bf16: flow rel-L2 {e:.2e}")


def test_a1_odd_tile_count_and_batches():
    """Shapes that exercise the pair-of-tiles scheduler: odd tile counts (N=20*16=320 -> 3 tiles), many items."""
    from emip_b200.matching import global_correlation_softmax
    for (b, h, w) in ((3, 20, 16), (5, 12, 12), (2, 45, 43)):
        f0 = cases.randn(70 + b, (b, 128, h, w), 1.5)
        f1 = cases.randn(80 + b, (b, 128, h, w), 1.5)
        ref_flow, _, ref_corr = O.global_correlation_softmax(f0, f1, True)
        flow, _, corr = global_correlation_softmax(dev(f0), dev(f1), True)
        assert rel(flow, ref_flow) < TOL_OUT, (b, h, w)
        assert rel(corr, ref_corr) < TOL_EXACT, (b, h, w)


@pytest.mark.parametrize("schedule", ["stream_k", "items"])
def test_a1_a2_work_schedules(schedule):
    """Both work schedules of the fused tcgen05 kernel (equal spans of key tiles with a cross-CTA merge of the
    online-softmax rows / whole items grid-strided) against the CPU oracle: small and c2-sized batches."""
    from emip_b200.matching import global_correlation_softmax
    from emip_b200.flow_attn import FeatureFlowAttention
    for (b, h, w) in ((1, 44, 44), (3, 20, 16), (11, 24, 20)):
        f0 = cases.randn(170 + b, (b, 128, h, w), 1.5)
        f1 = cases.randn(180 + b, (b, 128, h, w), 1.5)
        ref_flow, _, ref_corr = O.global_correlation_softmax(f0, f1, True)
        flow, _, corr = global_correlation_softmax(dev(f0), dev(f1), True, schedule=schedule)
        assert rel(flow, ref_flow) < TOL_OUT, (b, h, w)
        assert rel(corr, ref_corr) < TOL_EXACT, (b, h, w)
    # c2 size (16 pairs, 44 x 44): against the oracle itself
    f0 = cases.randn(2, (16, 128, 44, 44), 4.1)
    f1 = cases.randn(3, (16, 128, 44, 44), 4.1)
    ref_flow, _, ref_corr = O.global_correlation_softmax(f0, f1, True)
    flow_tc, _, corr_tc = global_correlation_softmax(dev(f0), dev(f1), True, schedule=schedule)
    assert rel(flow_tc, ref_flow) < TOL_OUT and rel(corr_tc, ref_corr) < TOL_EXACT
    # a2 (value table mode) against its golden vector
    name = list(cases.A2_CASES)[0]
    sp = cases.A2_CASES[name]
    d = cases.a2_inputs(sp)
    m = FeatureFlowAttention(sp["c"]).cuda()
    m.schedule = schedule
    m.load_state_dict({k: d[k] for k in ("q_proj.weight", "q_proj.bias", "k_proj.weight", "k_proj.bias")})
    ref = O.feature_flow_attention(d["x"], d["flow"], d["q_proj.weight"], d["q_proj.bias"], d["k_proj.weight"],
                                   d["k_proj.bias"])
    assert rel(m(dev(d["x"]), dev(d["flow"])), ref) < TOL_OUT


@pytest.mark.parametrize("exact", [True, False], ids=["exact_fp32", "tcgen05"])
@pytest.mark.parametrize("name", list(cases.A2_CASES))
def test_a2_flow_attention_golden(golden, name, exact):
    from emip_b200.flow_attn import FeatureFlowAttention
    g = golden(name)
    s = cases.A2_CASES[name]
    d = cases.a2_inputs(s)
    m = FeatureFlowAttention(s["c"]).cuda()
    m.load_state_dict({k: d[k] for k in ("q_proj.weight", "q_proj.bias", "k_proj.weight", "k_proj.bias")})
    m.exact_fp32 = exact
    x = dev(d["x"]).requires_grad_(True)
    out = m(x, dev(d["flow"]))
    e = cases.check_packed(out, g["out"], TOL_EXACT if exact else TOL_OUT, "out")
    print(f"{name} exact={exact}: out rel-L2 {e:.2e}")
    (out * dev(d["wout"])).sum().backward()
    cases.check_packed(x.grad, g["dx"], TOL_GRAD, "dx")
    cases.check_packed(m.q_proj.weight.grad, g["dqw"], TOL_GRAD, "dqw")
    cases.check_packed(m.q_proj.bias.grad, g["dqb"], TOL_GRAD, "dqb")
    cases.check_packed(m.k_proj.weight.grad, g["dkw"], TOL_GRAD, "dkw")


def test_a2_insitu_c1(golden):
    from emip_b200.flow_attn import FeatureFlowAttention
    g = golden("c1_insitu")
    m = FeatureFlowAttention(128).cuda()
    m.load_state_dict(g["ffa_params"])
    flow = dev(g["flow"]["full"])
    out = m(torch.cat((dev(g["f0"]), dev(g["f1"])), 0), flow)
    e = cases.check_packed(out, g["ffa_out"], TOL_OUT, "ffa_out")
    print(f"c1 in-situ FFA: rel-L2 {e:.2e}")


# ----------------------------------------------------------------------------- a4
@pytest.mark.parametrize("name", list(cases.A4_CASES))
@pytest.mark.parametrize("exact", [True, False], ids=["exact_fp32", "tcgen05"])
def test_a4_injector_golden(golden, name, exact):
    """Prompt fusion (feeder / collector) forward + full backward vs the reference's recorded outputs."""
    from emip_b200.injector import Injector
    g = golden(name)
    d = cases.a4_inputs(cases.A4_CASES[name])
    m = Injector().cuda()
    m.transformer.exact_fp32 = exact
    m.transformer.load_state_dict(d["params"])
    x = dev(d["x"]).requires_grad_(True)
    x1 = dev(d["x1"]).requires_grad_(True)
    out = m(x, x1)
    e = cases.check_packed(out, g["out"], 2e-5, "out")
    (out * dev(d["wout"])).sum().backward()
    e_dx = cases.check_packed(x.grad, g["dx"], 1e-4, "dx")
    e_dx1 = cases.check_packed(x1.grad, g["dx1"], 2e-4, "dx1")
    print(f"{name}: out {e:.2e} dx {e_dx:.2e} dx1 {e_dx1:.2e}")
    for k, p in m.transformer.named_parameters():
        cases.check_packed(p.grad, g["dparams"][k], 3e-4, "d" + k)


def test_a4_injector_vs_oracle_batch():
    """B=3 at full resolution vs the CPU oracle (forward), plus determinism of the two-stage reductions."""
    from emip_b200.injector import Injector
    s = dict(b=3, h=44, w=44, xs=2.2, x1s=1.0, seed=47)
    d = cases.a4_inputs(s)
    m = Injector().cuda()
    m.transformer.load_state_dict(d["params"])
    ref = O.injector(d["x"], d["x1"], d["params"])
    for exact in (True, False):
        m.transformer.exact_fp32 = exact
        out = m(dev(d["x"]), dev(d["x1"]))
        e = rel(out, ref)
        print(f"a4 B=3 exact={exact}: out rel-L2 {e:.2e}")
        assert e < 2e-5
        assert torch.equal(out, m(dev(d["x"]), dev(d["x1"])))


# ----------------------------------------------------------------------------- a5
@pytest.mark.parametrize("exact", [True, False], ids=["exact_fp32", "tcgen05"])
@pytest.mark.parametrize("name", list(cases.A5_CASES))
def test_a5_memory_read_golden(golden, name, exact):
    from emip_b200.memory import Memory
    g = golden(name)
    d = cases.a5_inputs(cases.A5_CASES[name])
    t = {k: dev(d[k]).requires_grad_(True) for k in ("m_in", "m_out", "q_in", "q_out")}
    mod = Memory()
    mod.exact_fp32 = exact
    out, p = mod(t["m_in"], t["m_out"], t["q_in"], t["q_out"])
    assert p is None
    e = cases.check_packed(out, g["out"], TOL_EXACT, "out")
    (out * dev(d["wout"])).sum().backward()
    errs = {k: cases.check_packed(v.grad, g["d" + k], 2e-4, "d" + k) for k, v in t.items()}
    print(f"{name}: out {e:.2e} grads {errs}")


@pytest.mark.parametrize("exact", [True, False], ids=["exact_fp32", "tcgen05"])
def test_a5_memory_read_t5_vs_oracle(exact):
    """The largest memory the model keeps (T = 5 frames, model_long.py:105-107), B = 1, inference path; and B = 3."""
    from emip_b200.memory import Memory
    mod = Memory()
    mod.exact_fp32 = exact
    for s in (dict(b=1, t=5, h=44, w=44, scale=1.5, seed=57), dict(b=3, t=2, h=20, w=28, scale=4.0, seed=58)):
        d = cases.a5_inputs(s)
        ref, _ = O.memory_read(d["m_in"], d["m_out"], d["q_in"], d["q_out"])
        with torch.no_grad():
            out, _ = mod(dev(d["m_in"]), dev(d["m_out"]), dev(d["q_in"]), dev(d["q_out"]))
        e = rel(out, ref)
        print(f"a5 exact={exact} {s}: rel-L2 {e:.2e}")
        assert e < TOL_EXACT


# ----------------------------------------------------------------------------- f4 (convex x8 upsampling)
@pytest.mark.parametrize("name", list(cases.F4_CASES))
def test_f4_convex_upsample_golden(golden, name):
    from emip_b200.upsample import upsample_flow_convex
    g = golden(name)
    d = cases.f4_inputs(cases.F4_CASES[name])
    flow = dev(d["flow"]).requires_grad_(True)
    mask = dev(d["mask"]).requires_grad_(True)
    out = upsample_flow_convex(flow, mask)
    e = cases.check_packed(out, g["out"], TOL_EXACT, "out")
    (out * dev(d["wout"])).sum().backward()
    e1 = cases.check_packed(flow.grad, g["dflow"], 1e-4, "dflow")
    e2 = cases.check_packed(mask.grad, g["dmask"], 1e-4, "dmask")
    print(f"{name}: out {e:.2e} dflow {e1:.2e} dmask {e2:.2e}")


def test_f4_convex_upsample_properties_batch():
    """B=32 (config c2's 2B): uniform mask = 3x3 box filter of 8*flow; deterministic backward."""
    from emip_b200.upsample import upsample_flow_convex
    B, h, w = 32, 44, 44
    g = torch.Generator(device="cuda").manual_seed(9)
    flow = 5 * torch.randn(B, 2, h, w, device="cuda", generator=g)
    out = upsample_flow_convex(flow, torch.zeros(B, 576, h, w, device="cuda"))
    box = torch.nn.functional.avg_pool2d(8 * flow, 3, 1, 1, count_include_pad=True)       # zero padding, /9
    assert rel(out, box.repeat_interleave(8, 2).repeat_interleave(8, 3)) < 1e-6
    mask = torch.randn(B, 576, h, w, device="cuda", generator=g).requires_grad_(True)
    f2 = flow.clone().requires_grad_(True)
    wout = torch.randn(B, 2, 8 * h, 8 * w, device="cuda", generator=g)
    upsample_flow_convex(f2, mask).backward(wout)
    g1 = (f2.grad.clone(), mask.grad.clone())
    f2.grad = None; mask.grad = None
    upsample_flow_convex(f2, mask).backward(wout)
    assert torch.equal(g1[0], f2.grad) and torch.equal(g1[1], mask.grad)


# ----------------------------------------------------------------------------- f3 (occlusion mask)
@pytest.mark.parametrize("name", list(cases.F3_CASES))
def test_f3_occlusion_mask_golden(golden, name):
    from emip_b200.warp import get_occu_mask_backward
    g = golden(name)
    fl4 = dev(cases.f3_inputs(cases.F3_CASES[name])["flow4"])
    mask = get_occu_mask_backward(fl4[:, 2:], th=0.2).cpu()
    ref, cm = g["mask"]["full"], g["corr_map"]["full"]
    assert mask.shape == ref.shape
    differ = (mask != ref) & ((cm - 0.2).abs() > 1e-5)              # only threshold ties may differ (atomic order)
    assert not differ.any(), int(differ.sum())
    assert 0.02 < mask.mean().item() < 0.9                          # the case exercises both classes


def test_f3_occlusion_mask_properties_full():
    """352x352, B=64: zero flow -> nothing occluded; a uniform shift occludes exactly the vacated border."""
    from emip_b200.warp import get_occu_mask_backward
    B, H, W = 64, 352, 352
    z = torch.zeros(B, 2, H, W, device="cuda")
    assert get_occu_mask_backward(z).sum().item() == 0
    z[:, 0] = 3.0                                                   # everything moves 3 px to the right
    m = get_occu_mask_backward(z)
    assert m[..., :3].min().item() == 1.0 and m[..., 3:].sum().item() == 0


# ----------------------------------------------------------------------------- f1
@pytest.mark.parametrize("name", list(cases.F1_CASES))
def test_f1_conv_corr_golden(golden, name):
    from emip_b200.conv_corr import conv_corr_first_layer
    g = golden(name)
    d = cases.f1_inputs(cases.F1_CASES[name])
    f0 = dev(d["f0"]).requires_grad_(True)
    f1 = dev(d["f1"]).requires_grad_(True)
    w = dev(d["weight"]).requires_grad_(True)
    b = dev(d["bias"]).requires_grad_(True)
    out = conv_corr_first_layer(f0, f1, w, b)
    e = cases.check_packed(out, g["out"], TOL_EXACT, "out")
    print(f"{name}: out rel-L2 {e:.2e}")
    (out * dev(d["wout"])).sum().backward()
    cases.check_packed(f0.grad, g["df0"], TOL_GRAD, "df0")
    cases.check_packed(f1.grad, g["df1"], TOL_GRAD, "df1")
    cases.check_packed(w.grad, g["dw"], TOL_GRAD, "dw")
    cases.check_packed(b.grad, g["db"], TOL_GRAD, "db")


def test_f1_conv_corr_model_size_vs_materialised():
    """The model's shape (44x44 tokens, 968 output channels): against F.conv2d on the cost volume our a1 kernel emits,
    weight update invalidates the prepared copy."""
    from emip_b200.conv_corr import conv_corr_first_layer
    from emip_b200.matching import global_correlation_softmax
    B, C, H, W, O = 2, 128, 44, 44, 968
    f0 = dev(cases.randn(95, (B, C, H, W), 4.1))
    f1 = dev(cases.randn(96, (B, C, H, W), 4.1))
    w = dev(cases.randn(97, (O, H * W, 3, 3), (9 * H * W) ** -0.5))
    b = dev(cases.randn(98, (O,), 0.1))
    corr = global_correlation_softmax(f0, f1, True, exact_fp32=True)[2]
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.functional.conv2d(corr.double(), w.double(), b.double(), padding=1)
    out = conv_corr_first_layer(f0, f1, w, b)
    assert rel(out, ref) < TOL_EXACT
    with torch.no_grad():
        w.mul_(0.5)
    assert rel(conv_corr_first_layer(f0, f1, w, b), (ref - b.double().view(1, -1, 1, 1)) * 0.5 + b.double().view(1, -1, 1, 1)) < TOL_EXACT


def test_f1_fused_model_slice():
    """dropin.fuse_conv_corr on a stand-in with the reference's structure (model.py:59-62,95-96): the fused forward and
    backward equal the unfused ones."""
    from emip_b200 import dropin
    from emip_b200.matching import global_correlation_softmax

    class Slice(torch.nn.Module):
        def __init__(self, n, o):
            super().__init__()
            self.conv_corr = torch.nn.Sequential(torch.nn.Conv2d(n, o, 3, 1, 1), torch.nn.BatchNorm2d(o),
                                                 torch.nn.ReLU(inplace=True), torch.nn.Conv2d(o, 16, 3, 1, 1))

        def forward(self, a, b):
            flow, _, corr = global_correlation_softmax(a, b, True)
            return flow, self.conv_corr(corr)

    B, C, H, W, O = 2, 128, 16, 16, 40
    torch.manual_seed(5)
    net = Slice(H * W, O).cuda()
    a = dev(cases.randn(31, (B, C, H, W), 1.5)).requires_grad_(True)
    b = dev(cases.randn(32, (B, C, H, W), 1.5)).requires_grad_(True)
    wf, wc = dev(cases.randn(33, (2 * B, 2, H, W))), dev(cases.randn(34, (B, 16, H, W)))

    def run():
        for t in (a, b):
            t.grad = None
        net.zero_grad()
        flow, y = net(a, b)
        ((flow * wf).sum() + (y * wc).sum()).backward()
        return flow.detach(), y.detach(), a.grad.clone(), b.grad.clone(), net.conv_corr[0].weight.grad.clone()

    ref = run()
    keys = list(net.state_dict())
    dropin.fuse_conv_corr(net)
    assert list(net.state_dict()) == keys
    got = run()
    for name, r, g in zip(("flow", "y", "da", "db", "dw"), ref, got):
        assert rel(g, r) < TOL_GRAD, name


# ----------------------------------------------------------------------------- f3b (photometric loss)
@pytest.mark.parametrize("name", list(cases.F3B_CASES))
def test_f3b_photometric_loss_golden(golden, name):
    from emip_b200.photometric import photometric_loss
    g = golden(name)
    d = cases.f3b_inputs(cases.F3B_CASES[name])
    rec = dev(d["rec"]).requires_grad_(True)
    loss = photometric_loss(dev(d["im"]), rec, dev(d["mask"]))
    assert abs(float(loss.detach()) - g["loss"]) <= 2e-6 * abs(g["loss"]), (float(loss.detach()), g["loss"])
    (3.0 * loss).backward()
    e = cases.check_packed(rec.grad / 3.0, g["drec"], 2e-5, "drec")
    print(f"{name}: loss {float(loss.detach()):.6f} vs {g['loss']:.6f}, drec rel-L2 {e:.2e}")


def test_f3b_photometric_loss_fullsize_with_warp():
    """The training call chain at full size (B reduced): flow_warp -> loss -> backward to the flow, vs the CPU oracle."""
    from emip_b200.photometric import photometric_loss
    from emip_b200.warp import flow_warp, get_occu_mask_backward
    B, H, W = 2, 352, 352
    im1, im2 = cases.randn(121, (B, 3, H, W), 0.5), cases.randn(122, (B, 3, H, W), 0.5)
    fl4 = torch.cat([cases.smooth_flow(123, B, H, W, 3.0), cases.smooth_flow(124, B, H, W, 3.0)], 1)
    f_ref = fl4.clone().requires_grad_(True)
    mask_ref = 1 - O.occu_mask_backward(fl4[:, 2:], 0.2)
    l_ref = O.photometric_loss(im1, O.flow_warp(im2, f_ref[:, :2]), mask_ref)
    l_ref.backward()
    f = dev(fl4).requires_grad_(True)
    mask = 1 - get_occu_mask_backward(f.detach()[:, 2:], th=0.2)
    loss = photometric_loss(dev(im1), flow_warp(dev(im2), f[:, :2]), mask)
    loss.backward()
    assert (mask.cpu() != mask_ref).float().mean() < 1e-4              # threshold ties only
    assert abs(float(loss.detach()) - float(l_ref.detach())) < 1e-4 * abs(float(l_ref.detach()))
    assert rel(f.grad[:, :2], f_ref.grad[:, :2]) < 2e-3


# ----------------------------------------------------------------------------- f2 (split-window attention)
@pytest.mark.parametrize("name", list(cases.F2_CASES))
def test_f2_split_window_attention_golden(golden, name):
    from emip_b200.window_attn import single_head_split_window_attention
    g = golden(name)
    s = cases.F2_CASES[name]
    d = cases.f2_inputs(s)
    for shift in (False, True):
        tag = "shift" if shift else "plain"
        q, k, v = (dev(d[n]).requires_grad_(True) for n in ("q", "k", "v"))
        mask = torch.zeros(1, device="cuda") if shift else None      # content implied by the geometry
        out = single_head_split_window_attention(q, k, v, num_splits=s["k"], with_shift=shift, h=s["h"], w=s["w"], attn_mask=mask)
        e = cases.check_packed(out, g[tag]["out"], TOL_EXACT, tag + " out")
        (out * dev(d["wout"])).sum().backward()
        cases.check_packed(q.grad, g[tag]["dq"], TOL_GRAD, tag + " dq")
        cases.check_packed(k.grad, g[tag]["dk"], TOL_GRAD, tag + " dk")
        cases.check_packed(v.grad, g[tag]["dv"], TOL_GRAD, tag + " dv")
        print(f"{name} {tag}: out rel-L2 {e:.2e}")


def test_f2_full_attention_vs_oracle():
    from emip_b200.window_attn import single_head_full_attention
    q, k, v = (cases.randn(140 + i, (3, 300, 128), 1.5) for i in range(3))
    ref = torch.softmax(q.double() @ k.double().transpose(1, 2) / 128 ** 0.5, -1) @ v.double()
    assert rel(single_head_full_attention(dev(q), dev(k), dev(v)), ref) < TOL_EXACT


@pytest.mark.parametrize("b,h,w,k", [(3, 44, 44, 2), (1, 24, 40, 2), (5, 12, 18, 3), (2, 16, 16, 1)])
def test_f2_window_geometries_vs_oracle(b, h, w, k):
    """The fused per-layer call against the oracle's roll + mask formulation on other batch sizes / grids / splits."""
    from emip_b200.window_attn import single_head_split_window_attention
    q, k_, v = (cases.randn(150 + i, (b, h * w, 128), 1.5 if i < 2 else 1.0) for i in range(3))
    for shift in (False, True):
        if shift and k == 1:
            continue                                                   # the reference never shifts a single window
        ref = O.split_window_attention(q.double(), k_.double(), v.double(), k, shift, h, w)
        out = single_head_split_window_attention(dev(q), dev(k_), dev(v), k, shift, h, w, torch.zeros(1) if shift else None)
        assert rel(out, ref) < TOL_EXACT, (shift, rel(out, ref))


def test_f2_attention_lazy_rescale_and_peaked_rows():
    """Row maxima that grow by > 2^16 from key tile to key tile (forces the O-accumulator rescale of attn_tc.cu on every
    tile) and near-one-hot rows with scores in [-80, 70]."""
    from emip_b200.window_attn import attention
    n = 700
    q = cases.randn(160, (2, n, 128))
    sgn = torch.sign(q[:, :1, :])
    k = cases.randn(161, (2, n, 128), 0.05) + torch.arange(n, dtype=torch.float32).view(1, n, 1) * 0.02 * sgn
    q = q.abs() * sgn
    v = cases.randn(162, (2, n, 128))
    s = q.double() @ k.double().transpose(1, 2) / 128 ** 0.5
    assert (s[:, :, 128:].amax(-1) - s[:, :, :128].amax(-1)).min() > 16 * 0.6931 * 3   # the ramp really crosses the threshold
    ref = torch.softmax(s, -1) @ v.double()
    assert rel(attention(dev(q), dev(k), dev(v)), ref) < 2e-4          # |S| ~ 160: S itself carries ~4e-6 * 160 of error
    q, k, v = (cases.randn(163 + i, (3, 200, 128), 4.0 if i < 2 else 1.0) for i in range(3))
    ref = torch.softmax(q.double() @ k.double().transpose(1, 2) / 128 ** 0.5, -1) @ v.double()
    assert rel(attention(dev(q), dev(k), dev(v)), ref) < TOL_EXACT


def test_f2_attention_empty_and_minimum():
    from emip_b200.window_attn import attention
    z = torch.zeros(0, 64, 128, device="cuda")
    assert attention(z, z, z).shape == (0, 64, 128)
    q, k, v = (cases.randn(170 + i, (2, 6, 128)) for i in range(3))     # a handful of tokens: one partial tile
    ref = torch.softmax(q.double() @ k.double().transpose(1, 2) / 128 ** 0.5, -1) @ v.double()
    assert rel(attention(dev(q), dev(k), dev(v)), ref) < TOL_EXACT


def test_a4_injector_odd_width_vs_oracle():
    """W % 4 != 0 takes the scalar depthwise / gate kernels (W % 4 == 0, every model size, takes the 4-wide ones):
    forward and all gradients against the CPU oracle."""
    from emip_b200.injector import Injector
    d = cases.a4_inputs(dict(b=2, h=9, w=10, xs=2.2, x1s=1.0, seed=49))
    m = Injector().cuda()
    m.transformer.load_state_dict(d["params"])
    x, x1 = d["x"].clone().requires_grad_(True), d["x1"].clone().requires_grad_(True)
    prm = {k: v.clone().requires_grad_(True) for k, v in d["params"].items()}
    ref = O.injector(x, x1, prm)
    (ref * d["wout"]).sum().backward()
    gx, gx1 = dev(d["x"]).requires_grad_(True), dev(d["x1"]).requires_grad_(True)
    out = m(gx, gx1)
    (out * dev(d["wout"])).sum().backward()
    assert rel(out, ref) < 2e-5
    assert rel(gx.grad, x.grad) < 3e-4 and rel(gx1.grad, x1.grad) < 3e-4
    for k, p in m.transformer.named_parameters():
        assert rel(p.grad, prm[k].grad) < 3e-4, k


@pytest.mark.parametrize("nb,n,scale", [(2, 100, 1.0), (3, 300, 1.5), (2, 700, 1.0), (5, 16, 2.0)])
def test_f2_attention_backward_tc_vs_fp64(nb, n, scale):
    """dq, dk, dv of the tensor-core backward (attn_bwd_tc.cu: three launches of one kernel) against fp64 autograd,
    incl. partial row / column tiles and several column tiles per row tile."""
    from emip_b200.window_attn import attention
    q, k, v, w = (cases.randn(180 + i, (nb, n, 128), scale if i < 2 else 1.0) for i in range(4))
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    (torch.softmax(qd @ kd.transpose(1, 2) / 128 ** 0.5, -1) @ vd).backward(w.double())
    qq, kk, vv = (dev(t).requires_grad_(True) for t in (q, k, v))
    attention(qq, kk, vv).backward(dev(w))
    for name, a, b in (("dq", qq.grad, qd.grad), ("dk", kk.grad, kd.grad), ("dv", vv.grad, vd.grad)):
        assert rel(a, b) < 1e-4, (name, rel(a, b))


def test_a5_memory_read_backward_tc_vs_exact():
    """a5 backward on the tensor cores (key-split row launch, CN layouts, strided dmem) against the exact-fp32 kernels."""
    from emip_b200.memory import Memory
    d = cases.a5_inputs(dict(b=2, t=3, h=20, w=24, scale=1.5, seed=59))
    grads = {}
    for exact in (True, False):
        t = {k: dev(d[k]).requires_grad_(True) for k in ("m_in", "m_out", "q_in", "q_out")}
        m = Memory()
        m.exact_fp32 = exact
        out, _ = m(t["m_in"], t["m_out"], t["q_in"], t["q_out"])
        (out * dev(d["wout"])).sum().backward()
        grads[exact] = {k: v.grad.clone() for k, v in t.items()}
    for k in ("m_in", "m_out", "q_in", "q_out"):
        assert rel(grads[False][k], grads[True][k]) < 1e-4, (k, rel(grads[False][k], grads[True][k]))


def test_f1_backward_tc_no_bias_single_sample_vs_fp64():
    """f1 backward (five tensor-core GEMMs) without a bias and with B = 1, against fp64 autograd through the
    materialised cost volume (the reference's formulation)."""
    from emip_b200.conv_corr import conv_corr_first_layer
    B, C, H, W, Oc = 1, 128, 12, 20, 72
    f0, f1 = cases.randn(201, (B, C, H, W), 2.0), cases.randn(202, (B, C, H, W), 2.0)
    w = cases.randn(203, (Oc, H * W, 3, 3), (9 * H * W) ** -0.5)
    wo = cases.randn(204, (B, Oc, H, W))
    a, b_, c = (t.double().requires_grad_(True) for t in (f0, f1, w))
    O.conv_corr_first_layer(a, b_, c, None).backward(wo.double())
    ga, gb, gc = (dev(t).requires_grad_(True) for t in (f0, f1, w))
    conv_corr_first_layer(ga, gb, gc, None).backward(dev(wo))
    for name, x, y in (("df0", ga.grad, a.grad), ("df1", gb.grad, b_.grad), ("dw", gc.grad, c.grad)):
        assert rel(x, y) < 1e-4, (name, rel(x, y))


# ----------------------------------------------------------------------------- f2b (token-major layers of the blocks)
class _Layer(torch.nn.Module):
    """Same submodules / state_dict keys as the reference TransformerLayer (transformer.py:126-148), frozen as train.py:340-342."""

    def __init__(self, params, no_ffn, with_shift, c=128):
        super().__init__()
        nn = torch.nn
        self.attention_type, self.nhead, self.no_ffn, self.with_shift = "swin", 1, no_ffn, with_shift
        self.q_proj, self.k_proj, self.v_proj, self.merge = (nn.Linear(c, c, bias=False) for _ in range(4))
        self.norm1 = nn.LayerNorm(c)
        if not no_ffn:
            self.mlp = nn.Sequential(nn.Linear(2 * c, 8 * c, bias=False), nn.GELU(), nn.Linear(8 * c, c, bias=False))
            self.norm2 = nn.LayerNorm(c)
        self.load_state_dict({k: v for k, v in params.items() if k in self.state_dict()})
        self.requires_grad_(False).cuda()


@pytest.mark.parametrize("name", list(cases.F2B_CASES))
def test_f2b_transformer_layer_golden(golden, name):
    from emip_b200.transformer_layer import transformer_layer_forward
    g = golden(name)
    s = cases.F2B_CASES[name]
    d = cases.f2b_inputs(s)
    mask = torch.zeros(1, device="cuda")
    for no_ffn in (True, False):
        for shift in (False, True):
            tag = ("self" if no_ffn else "cross") + ("_shift" if shift else "_plain")
            layer = _Layer(d["params"], no_ffn, shift)
            src, tgt = dev(d["source"]).requires_grad_(True), dev(d["target"]).requires_grad_(True)
            out = transformer_layer_forward(layer, src, tgt, height=s["h"], width=s["w"], shifted_window_attn_mask=mask,
                                            attn_num_splits=s["k"])
            e = cases.check_packed(out, g[tag]["out"], TOL_EXACT, tag + " out")
            (out * dev(d["wout"])).sum().backward()
            cases.check_packed(src.grad, g[tag]["dsource"], TOL_GRAD, tag + " dsource")
            cases.check_packed(tgt.grad, g[tag]["dtarget"], TOL_GRAD, tag + " dtarget")
            print(f"{name} {tag}: out rel-L2 {e:.2e}")


@pytest.mark.parametrize("L,M,K,gelu", [(1936, 128, 128, False), (777, 1024, 256, False), (1000, 128, 1024, True), (5, 36, 20, True)])
def test_f2b_linear_tm_vs_fp64(L, M, K, gelu):
    """Token-major bias-free linear layer (ragged row counts, non-multiple-of-64 widths, GELU on load), forward and
    input gradient, against fp64."""
    from emip_b200.transformer_layer import linear_tm
    x, w, wo = cases.randn(301, (L, K), 1.5), cases.randn(302, (M, K), K ** -0.5), cases.randn(303, (L, M))
    xd = x.double().requires_grad_(True)
    ref = (O._gelu_erf(xd) if gelu else xd) @ w.double().T
    ref.backward(wo.double())
    xg = dev(x).requires_grad_(True)
    out = linear_tm(xg, dev(w), gelu_in=gelu)
    out.backward(dev(wo))
    assert rel(out, ref) < TOL_EXACT, rel(out, ref)
    assert rel(xg.grad, xd.grad) < TOL_EXACT, rel(xg.grad, xd.grad)


def test_f2b_layer_norm_tm_vs_fp64_and_errors():
    from emip_b200.transformer_layer import layer_norm_tm, linear_tm
    from emip_b200._lib import EmipError
    L = 1001
    x, r, g, b, wo = (cases.randn(311 + i, shp) for i, shp in enumerate(((L, 128), (L, 128), (128,), (128,), (L, 128))))
    x = x * 3 + 0.5
    xd, rd = x.double().requires_grad_(True), r.double().requires_grad_(True)
    ref = rd + torch.nn.functional.layer_norm(xd, (128,), g.double(), b.double(), 1e-5)
    ref.backward(wo.double())
    xg, rg = dev(x).requires_grad_(True), dev(r).requires_grad_(True)
    out = layer_norm_tm(xg, dev(g), dev(b), 1e-5, residual=rg)
    out.backward(dev(wo))
    assert rel(out, ref) < 1e-6 and rel(xg.grad, xd.grad) < 1e-5 and torch.equal(rg.grad.cpu(), wo)
    assert layer_norm_tm(dev(x)[:0], dev(g), dev(b)).shape == (0, 128)
    with pytest.raises(EmipError):
        layer_norm_tm(dev(x)[:, :64].contiguous(), dev(g)[:64], dev(b)[:64])         # only C = 128 is built
    with pytest.raises(EmipError):
        linear_tm(x, g.view(1, 128))                                                  # CPU tensors: no fallback
    with pytest.raises(NotImplementedError):
        linear_tm(dev(x), dev(r)[:128].clone().requires_grad_(True))                  # frozen weights only


@pytest.mark.parametrize("L,K1,Hd,M", [(1936, 256, 1024, 128), (333, 40, 192, 36)])
def test_f2b_mlp_tm_fused_vs_fp64(L, K1, Hd, M):
    """The no-grad MLP call (GELU + hi | lo split in the first GEMM's epilogue, pre-split A operand of the second)
    against fp64 and against the two-call path."""
    from emip_b200.transformer_layer import mlp_tm, linear_tm
    x, w1, w2 = cases.randn(321, (L, K1), 1.5), cases.randn(322, (Hd, K1), K1 ** -0.5), cases.randn(323, (M, Hd), Hd ** -0.5)
    ref = O._gelu_erf(x.double() @ w1.double().T) @ w2.double().T
    with torch.no_grad():
        out = mlp_tm(dev(x), dev(w1), dev(w2))
        two = linear_tm(linear_tm(dev(x), dev(w1)), dev(w2), gelu_in=True)
    assert rel(out, ref) < TOL_EXACT, rel(out, ref)
    assert rel(out, two) < TOL_EXACT, rel(out, two)
    assert mlp_tm(dev(x)[:0], dev(w1), dev(w2)).shape == (0, M)
    # autograd path: pre-activation kept, GELU' applied inside the operand split of the input-gradient GEMM
    from emip_b200.transformer_layer import _MlpTM
    wo = cases.randn(324, (L, M))
    xd = x.double().requires_grad_(True)
    (O._gelu_erf(xd @ w1.double().T) @ w2.double().T).backward(wo.double())
    xg = dev(x).requires_grad_(True)
    og = _MlpTM.apply(xg, dev(w1), dev(w2))
    og.backward(dev(wo))
    assert rel(og, ref) < TOL_EXACT and rel(xg.grad, xd.grad) < TOL_EXACT, (rel(og, ref), rel(xg.grad, xd.grad))


@pytest.mark.parametrize("name", list(cases.F2B_CASES))
def test_f2b_transformer_layer_inference_path_golden(golden, name):
    """The no-grad path (LayerNorm + residual in the GEMM epilogues, fused MLP call) against the reference's vectors, and
    against the autograd path of the same layer."""
    from emip_b200.transformer_layer import transformer_layer_forward, linear_ln_tm, linear_tm, layer_norm_tm
    g = golden(name)
    s = cases.F2B_CASES[name]
    d = cases.f2b_inputs(s)
    mask = torch.zeros(1, device="cuda")
    kw = dict(height=s["h"], width=s["w"], shifted_window_attn_mask=mask, attn_num_splits=s["k"])
    for no_ffn in (True, False):
        for shift in (False, True):
            tag = ("self" if no_ffn else "cross") + ("_shift" if shift else "_plain")
            layer = _Layer(d["params"], no_ffn, shift)
            with torch.no_grad():
                out = transformer_layer_forward(layer, dev(d["source"]), dev(d["target"]), **kw)
            cases.check_packed(out, g[tag]["out"], TOL_EXACT, tag + " out (no-grad path)")
            ref = transformer_layer_forward(layer, dev(d["source"]).requires_grad_(True), dev(d["target"]), **kw)
            assert rel(out, ref) < 1e-5, (tag, rel(out, ref))
    # ragged row count, offset mean, no residual
    x, w = cases.randn(331, (777, 128), 2.0) + 0.7, cases.randn(332, (128, 128), 128 ** -0.5)
    gm, bt = 1 + 0.2 * cases.randn(333, (128,)), 0.3 * cases.randn(334, (128,))
    with torch.no_grad():
        a = linear_ln_tm(dev(x), dev(w), dev(gm), dev(bt), 1e-5)
        b = layer_norm_tm(linear_tm(dev(x), dev(w)), dev(gm), dev(bt), 1e-5)
    refd = torch.nn.functional.layer_norm(x.double() @ w.double().T, (128,), gm.double(), bt.double(), 1e-5)
    assert rel(a, refd) < TOL_EXACT and rel(a, b) < 1e-5, (rel(a, refd), rel(a, b))


def test_f2b_self_layer_shared_split_vs_separate_calls():
    """source is target (how TransformerBlock calls the self-attention layer, transformer.py:383-389): q / k / v come from one
    operand split (emip_linear_tm_multi_fwd); same outputs and gradients as the three separate calls, and as fp64."""
    from emip_b200.transformer_layer import transformer_layer_forward, linear_tm_multi
    s = cases.F2B_CASES["f2b_small"]
    d = cases.f2b_inputs(s)
    layer = _Layer(d["params"], True, True)
    kw = dict(height=s["h"], width=s["w"], shifted_window_attn_mask=torch.zeros(1, device="cuda"), attn_num_splits=s["k"])
    x = dev(d["source"]).requires_grad_(True)
    a = transformer_layer_forward(layer, x, x, **kw)
    a.backward(dev(d["wout"]))
    x1, x2 = dev(d["source"]).requires_grad_(True), dev(d["source"]).requires_grad_(True)
    b = transformer_layer_forward(layer, x1, x2, **kw)
    b.backward(dev(d["wout"]))
    assert rel(a, b) < 1e-6 and rel(x.grad, x1.grad + x2.grad) < 1e-5, (rel(a, b), rel(x.grad, x1.grad + x2.grad))
    with torch.no_grad():
        assert rel(transformer_layer_forward(layer, x.detach(), x.detach(), **kw), b) < 1e-5
    xs, ws = cases.randn(341, (301, 192), 1.3), [cases.randn(342 + i, (40, 192), 192 ** -0.5) for i in range(3)]
    ys = linear_tm_multi(dev(xs), [dev(w) for w in ws])
    for y, w in zip(ys, ws):
        assert rel(y, xs.double() @ w.double().T) < TOL_EXACT


# ------------------------------------------------------------------- BASELINE config sizes against the oracle itself
def test_a1_c2_vs_oracle():
    """c2 (B = 16, 44 x 44, C = 128, bidirectional, corr): the tensor-core path against the CPU oracle, forward + backward."""
    from emip_b200.matching import global_correlation_softmax
    f0 = cases.randn(2, (16, 128, 44, 44), 4.1)
    f1 = cases.randn(3, (16, 128, 44, 44), 4.1)
    wf = cases.randn(4, (32, 2, 44, 44))
    c0, c1 = f0.clone().requires_grad_(True), f1.clone().requires_grad_(True)
    rflow, _, rcorr = O.global_correlation_softmax(c0, c1, True)
    (rflow * wf).sum().backward()
    g0, g1 = dev(f0).requires_grad_(True), dev(f1).requires_grad_(True)
    flow, _, corr = global_correlation_softmax(g0, g1, True)
    (flow * dev(wf)).sum().backward()
    assert rel(flow, rflow) < TOL_OUT and rel(corr, rcorr) < TOL_EXACT
    assert rel(g0.grad, c0.grad) < TOL_GRAD and rel(g1.grad, c1.grad) < TOL_GRAD


def test_a1_a2_c3_batch64_vs_oracle():
    """c3 / c5 size (B = 64 pairs): one call on the GPU, the oracle on chunks of 16 samples of the same batch."""
    from emip_b200.matching import global_correlation_softmax
    from emip_b200.flow_attn import FeatureFlowAttention
    f0 = cases.randn(5, (64, 128, 44, 44), 4.1)
    f1 = cases.randn(6, (64, 128, 44, 44), 4.1)
    flow, _, _ = global_correlation_softmax(dev(f0), dev(f1), True, return_corr=False)
    prm = cases.ffa_params(5)
    m = FeatureFlowAttention(128).cuda()
    m.load_state_dict(prm)
    with torch.no_grad():
        prop = m(torch.cat((dev(f0), dev(f1)), 0), flow)
    for c in range(0, 64, 16):
        sl = slice(c, c + 16)
        rflow, _, _ = O.global_correlation_softmax(f0[sl], f1[sl], True)
        assert rel(flow[sl], rflow[:16]) < TOL_OUT and rel(flow[64:][sl], rflow[16:]) < TOL_OUT, c
        if c == 16:       # flow propagation (2B = 128 maps on the GPU) on one chunk: forward flows of f0, backward flows of f1
            x = torch.cat((f0[sl], f1[sl]), 0)
            ref = O.feature_flow_attention(x, rflow, prm["q_proj.weight"], prm["q_proj.bias"], prm["k_proj.weight"], prm["k_proj.bias"])
            assert rel(prop[sl], ref[:16]) < TOL_OUT and rel(prop[64:][sl], ref[16:]) < TOL_OUT


def test_a2_c2_batch32_vs_oracle():
    """a2 at 2B = 32 maps (the c2 batch) against the oracle, forward + backward to the features."""
    from emip_b200.flow_attn import FeatureFlowAttention
    x = cases.randn(7, (32, 128, 44, 44), 4.1)
    fl = cases.randn(8, (32, 2, 44, 44), 12.0)
    wo = cases.randn(9, (32, 2, 44, 44))
    prm = cases.ffa_params(6)
    m = FeatureFlowAttention(128).cuda()
    m.load_state_dict(prm)
    cx = x.clone().requires_grad_(True)
    ref = O.feature_flow_attention(cx, fl, prm["q_proj.weight"], prm["q_proj.bias"], prm["k_proj.weight"], prm["k_proj.bias"])
    (ref * wo).sum().backward()
    gx = dev(x).requires_grad_(True)
    out = m(gx, dev(fl))
    (out * dev(wo)).sum().backward()
    assert rel(out, ref) < TOL_OUT and rel(gx.grad, cx.grad) < TOL_GRAD


@pytest.mark.parametrize("name", list(cases.A2_CASES))
def test_a2_bf16_mode(golden, name):
    """FeatureFlowAttention.bf16 = True (c3's single-pass bf16 operand mode): within 2e-2 of the fp32 reference."""
    from emip_b200.flow_attn import FeatureFlowAttention
    g = golden(name)
    s = cases.A2_CASES[name]
    d = cases.a2_inputs(s)
    m = FeatureFlowAttention(s["c"]).cuda()
    m.load_state_dict({k: d[k] for k in ("q_proj.weight", "q_proj.bias", "k_proj.weight", "k_proj.bias")})
    m.bf16 = True
    with torch.no_grad():
        out = m(dev(d["x"]), dev(d["flow"]))
    e = cases.check_packed(out, g["out"], TOL_BF16, "out")
    print(f"{name} a2 bf16: rel-L2 {e:.2e}")
    x = dev(d["x"]).requires_grad_(True)
    with pytest.raises(Exception):                  # forward-only mode: the backward refuses
        m(x, dev(d["flow"])).sum().backward()


def test_a4_batch64_subsampled_vs_oracle():
    """a4 at the c3 batch (64 samples in one call): four samples of the batch against the oracle."""
    from emip_b200.injector import injector_forward
    p = cases.injector_params(44)
    x = cases.randn(441, (64, 128, 44, 44), 2.2)
    x1 = cases.randn(442, (64, 128, 44, 44), 1.0)
    with torch.no_grad():
        out = injector_forward(dev(x), dev(x1), {k: dev(v) for k, v in p.items()})
    idx = [0, 21, 42, 63]
    ref = O.injector(x[idx], x1[idx], p)
    assert rel(out[idx], ref) < 1e-4


def test_a3_c5_batch64_vs_oracle():
    """a3 at B = 64, 3 x 352 x 352 (config c5's call): two samples of the batch against the oracle, both flow slices."""
    from emip_b200.warp import flow_warp
    B, H, W = 64, 352, 352
    x = cases.randn(91, (B, 3, H, W))
    fl4 = torch.cat([cases.smooth_flow(92, B, H, W, 12.0), cases.smooth_flow(93, B, H, W, 12.0)], 1)
    xg, fg = dev(x), dev(fl4)
    for sl in (slice(0, 2), slice(2, 4)):
        out = flow_warp(xg, fg[:, sl])
        for b in (5, 63):
            ref = O.flow_warp(x[b:b + 1], fl4[b:b + 1, sl])
            assert rel(out[b:b + 1], ref) < 1e-5


def test_a3_two_threads_two_streams():
    """Forward on one thread / stream and backward on another at the same time (the autograd engine does exactly that):
    the library keeps no kernel-choice state, the host-side choice is lock-guarded; results stay bit-identical."""
    import threading
    from emip_b200.warp import flow_warp, KERNEL_DIRECT, KERNEL_STAGED
    B, H, W = 8, 352, 352
    x = dev(cases.randn(95, (B, 3, H, W)))
    smooth = dev(cases.smooth_flow(96, B, H, W, 6.0))
    noisy = dev(cases.randn(97, (B, 2, H, W), 20.0))
    wout = dev(cases.randn(98, (B, 3, H, W)))
    ref = {}
    for name, fl in (("smooth", smooth), ("noisy", noisy)):
        f = fl.clone().requires_grad_(True)
        o = flow_warp(x, f, kernel=KERNEL_DIRECT)
        (o * wout).sum().backward()
        ref[name] = (o.detach().clone(), f.grad.clone())
        f2 = fl.clone().requires_grad_(True)
        o2 = flow_warp(x, f2, kernel=KERNEL_STAGED)
        (o2 * wout).sum().backward()
        assert torch.equal(o2, ref[name][0]) and torch.equal(f2.grad, ref[name][1])     # the two kernels are bit-identical
    torch.cuda.synchronize()
    errors = []

    def worker(name, fl, n):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(n):
                    f = fl.clone().requires_grad_(True)
                    o = flow_warp(x, f)                    # adaptive choice, shared per-device state
                    (o * wout).sum().backward()
                    if not (torch.equal(o, ref[name][0]) and torch.equal(f.grad, ref[name][1])):
                        errors.append(name)
            st.synchronize()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    ts = [threading.Thread(target=worker, args=("smooth", smooth, 60)), threading.Thread(target=worker, args=("noisy", noisy, 60))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:3]


def test_bf16_modes_are_forward_only_and_memory_q_out_only_grad():
    """the bf16 inference modes refuse a backward pass (their saved log-sum-exp comes from bf16 scores); a memory read
    whose only differentiable input is q_out returns the pass-through half of the incoming gradient."""
    from emip_b200.matching import global_correlation_softmax
    from emip_b200.memory import Memory
    f0 = dev(cases.randn(501, (1, 128, 8, 8))).requires_grad_(True)
    f1 = dev(cases.randn(502, (1, 128, 8, 8)))
    flow, _, _ = global_correlation_softmax(f0, f1, True, bf16=True)
    with pytest.raises(Exception):
        flow.sum().backward()
    d5 = cases.a5_inputs(dict(b=1, t=2, h=5, w=6, scale=1.0, seed=58))
    q_out = dev(d5["q_out"]).requires_grad_(True)
    out, _ = Memory()(dev(d5["m_in"]), dev(d5["m_out"]), dev(d5["q_in"]), q_out)
    wout = dev(d5["wout"])
    (out * wout).sum().backward()
    assert torch.equal(q_out.grad.reshape(1, 128, 5, 6), wout[:, 128:])
