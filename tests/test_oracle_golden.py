"""Pin the CPU oracle (oracle/restate.py) against golden vectors recorded from the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

import cases
from oracle import restate as O

TOL = 2e-5      # fp32 re-association only
GTOL = 1e-4     # gradients through near-one-hot softmaxes


@pytest.mark.parametrize("name", list(cases.A1_CASES))
def test_a1_global_matching(golden, name):
    g = golden(name)
    s = cases.A1_CASES[name]
    d = cases.a1_inputs(s)
    f0 = d["f0"].clone().requires_grad_(True)
    f1 = d["f1"].clone().requires_grad_(True)
    flow, prob, corr = O.global_correlation_softmax(f0, f1, True)
    assert prob is None and tuple(corr.shape) == (s["b"], s["h"] * s["w"], s["h"], s["w"])
    cases.check_packed(flow, g["flow"], TOL, "flow")
    cases.check_packed(corr, g["corr"], TOL, "corr")
    ((flow * d["wflow"]).sum() + (corr * d["wcorr"]).sum()).backward()
    cases.check_packed(f0.grad, g["df0"], GTOL, "df0")
    cases.check_packed(f1.grad, g["df1"], GTOL, "df1")


def test_a1_unidirectional_and_prob():
    s = cases.A1_CASES["a1_small"]
    d = cases.a1_inputs(s)
    flow, prob, _ = O.global_correlation_softmax(d["f0"], d["f1"], False, return_prob=True)
    flow2, _, _ = O.global_correlation_softmax(d["f0"], d["f1"], True)
    assert flow.shape[0] == s["b"] and torch.allclose(flow, flow2[: s["b"]], atol=1e-5)
    assert (prob.sum(-1) - 1).abs().max() < 1e-5


def test_a1_insitu_c1(golden):
    g = golden("c1_insitu")
    flow, _, corr = O.global_correlation_softmax(g["f0"], g["f1"], True)
    cases.check_packed(flow, g["flow"], TOL, "flow")
    cases.check_packed(corr, g["corr"], TOL, "corr")
    p = g["ffa_params"]
    out = O.feature_flow_attention(torch.cat((g["f0"], g["f1"]), 0), flow, p["q_proj.weight"], p["q_proj.bias"],
                                   p["k_proj.weight"], p["k_proj.bias"])
    cases.check_packed(out, g["ffa_out"], 5e-5, "ffa_out")


@pytest.mark.parametrize("name", list(cases.A2_CASES))
def test_a2_flow_attention(golden, name):
    g = golden(name)
    d = cases.a2_inputs(cases.A2_CASES[name])
    x = d["x"].clone().requires_grad_(True)
    qw = d["q_proj.weight"].clone().requires_grad_(True)
    qb = d["q_proj.bias"].clone().requires_grad_(True)
    kw = d["k_proj.weight"].clone().requires_grad_(True)
    out = O.feature_flow_attention(x, d["flow"], qw, qb, kw, d["k_proj.bias"])
    cases.check_packed(out, g["out"], TOL, "out")
    (out * d["wout"]).sum().backward()
    cases.check_packed(x.grad, g["dx"], GTOL, "dx")
    cases.check_packed(qw.grad, g["dqw"], GTOL, "dqw")
    cases.check_packed(qb.grad, g["dqb"], GTOL, "dqb")
    cases.check_packed(kw.grad, g["dkw"], GTOL, "dkw")


@pytest.mark.parametrize("name", list(cases.A3_CASES))
def test_a3_flow_warp(golden, name):
    g = golden(name)
    s = cases.A3_CASES[name]
    d = cases.a3_inputs(s)
    x = d["x"].clone().requires_grad_(True)
    fl = d["flow"].clone().requires_grad_(True)
    out = O.flow_warp(x, fl, pad=s["pad"])
    cases.check_packed(out, g["out"], 1e-5, "out")
    (out * d["wout"]).sum().backward()
    cases.check_packed(fl.grad, g["dflow"], 1e-4, "dflow")
    cases.check_packed(x.grad, g["dx"], 1e-5, "dx")


def test_a3_noncontiguous_slice(golden):
    g = golden("a3_slice")
    s = g["spec"]
    flow4 = cases.randn(s["seed"], (s["b"], 4, s["h"], s["w"]), s["sigma"])
    x = cases.randn(s["seed"] + 1, (s["b"], 3, s["h"], s["w"]))
    cases.check_packed(O.flow_warp(x, flow4[:, :2]), g["out_fw"], 1e-5)
    cases.check_packed(O.flow_warp(x, flow4[:, 2:]), g["out_bw"], 1e-5)


def test_a3_matches_aten_grid_sample():
    """The published ATen semantics the restatement follows, checked against the installed torch."""
    import torch.nn.functional as F
    x = cases.randn(5, (2, 3, 16, 20))
    fl = cases.randn(6, (2, 2, 16, 20), 6.0)
    for pad in ("border", "zeros"):
        h, w = 16, 20
        gx = torch.arange(w).view(1, 1, w) + fl[:, 0]
        gy = torch.arange(h).view(1, h, 1) + fl[:, 1]
        grid = torch.stack([2.0 * gx / (w - 1) - 1.0, 2.0 * gy / (h - 1) - 1.0], dim=-1)
        ref = F.grid_sample(x, grid, mode="bilinear", padding_mode=pad, align_corners=True)
        assert (O.flow_warp(x, fl, pad) - ref).abs().max() < 2e-6


@pytest.mark.parametrize("name", list(cases.A4_CASES))
def test_a4_injector(golden, name):
    g = golden(name)
    d = cases.a4_inputs(cases.A4_CASES[name])
    p = {k: v.clone().requires_grad_(True) for k, v in d["params"].items()}
    x = d["x"].clone().requires_grad_(True)
    x1 = d["x1"].clone().requires_grad_(True)
    out = O.injector(x, x1, p)
    cases.check_packed(out, g["out"], TOL, "out")
    (out * d["wout"]).sum().backward()
    cases.check_packed(x.grad, g["dx"], GTOL, "dx")
    cases.check_packed(x1.grad, g["dx1"], GTOL, "dx1")
    for k in p:
        cases.check_packed(p[k].grad, g["dparams"][k], 2e-4, "d" + k)


@pytest.mark.parametrize("name", list(cases.A5_CASES))
def test_a5_memory_read(golden, name):
    g = golden(name)
    d = cases.a5_inputs(cases.A5_CASES[name])
    t = {k: d[k].clone().requires_grad_(True) for k in ("m_in", "m_out", "q_in", "q_out")}
    out, p = O.memory_read(t["m_in"], t["m_out"], t["q_in"], t["q_out"])
    assert p is None
    cases.check_packed(out, g["out"], TOL, "out")
    (out * d["wout"]).sum().backward()
    for k, v in t.items():
        cases.check_packed(v.grad, g["d" + k], GTOL, "d" + k)


@pytest.mark.parametrize("name", list(cases.F4_CASES))
def test_f4_convex_upsample(golden, name):
    g = golden(name)
    d = cases.f4_inputs(cases.F4_CASES[name])
    flow = d["flow"].clone().requires_grad_(True)
    mask = d["mask"].clone().requires_grad_(True)
    out = O.upsample_flow_convex(flow, mask)
    cases.check_packed(out, g["out"], TOL, "out")
    (out * d["wout"]).sum().backward()
    cases.check_packed(flow.grad, g["dflow"], GTOL, "dflow")
    cases.check_packed(mask.grad, g["dmask"], GTOL, "dmask")


@pytest.mark.parametrize("name", list(cases.F3_CASES))
def test_f3_occlusion_mask(golden, name):
    g = golden(name)
    fl = cases.f3_inputs(cases.F3_CASES[name])["flow4"][:, 2:]
    cm = O.corresponding_map(fl)
    cases.check_packed(cm, g["corr_map"], TOL, "corr_map")
    mask = O.occu_mask_backward(fl, 0.2)
    ref = g["mask"]["full"]
    differ = (mask != ref) & ((cm - 0.2).abs() > 1e-5)              # only threshold ties may differ (summation order)
    assert not differ.any()


@pytest.mark.parametrize("name", list(cases.F1_CASES))
def test_f1_conv_corr_first_layer(golden, name):
    """conv_corr[0] on the cost volume: the oracle restatement and the two-GEMM re-association the CUDA path uses."""
    g = golden(name)
    d = cases.f1_inputs(cases.F1_CASES[name])
    cases.check_packed(O.conv_corr_first_layer(d["f0"], d["f1"], d["weight"], d["bias"]), g["out"], TOL, "out")
    f0 = d["f0"].clone().requires_grad_(True)
    f1 = d["f1"].clone().requires_grad_(True)
    w = d["weight"].clone().requires_grad_(True)
    b = d["bias"].clone().requires_grad_(True)
    out = O.conv_corr_reassociated(f0, f1, w, b)
    cases.check_packed(out, g["out"], 1e-5, "re-associated out")
    (out * d["wout"]).sum().backward()
    cases.check_packed(f0.grad, g["df0"], 1e-5, "df0")
    cases.check_packed(f1.grad, g["df1"], 1e-5, "df1")
    cases.check_packed(w.grad, g["dw"], 1e-5, "dw")
    cases.check_packed(b.grad, g["db"], 1e-5, "db")


@pytest.mark.parametrize("name", list(cases.F3B_CASES))
def test_f3b_photometric_loss(golden, name):
    g = golden(name)
    d = cases.f3b_inputs(cases.F3B_CASES[name])
    rec = d["rec"].clone().requires_grad_(True)
    loss = O.photometric_loss(d["im"], rec, d["mask"])
    assert abs(float(loss.detach()) - g["loss"]) <= 1e-6 * abs(g["loss"])
    loss.backward()
    cases.check_packed(rec.grad, g["drec"], 1e-5, "drec")


@pytest.mark.parametrize("name", list(cases.F2_CASES))
def test_f2_split_window_attention(golden, name):
    """The oracle restatement of the swin split-window attention, and the block decomposition the CUDA path uses for
    the shifted layers (plain attention inside the rectangles the 0 / -100 mask separates)."""
    from emip_b200.window_attn import _axis_groups
    g = golden(name)
    s = cases.F2_CASES[name]
    d = cases.f2_inputs(s)
    for shift in (False, True):
        tag = "shift" if shift else "plain"
        out = O.split_window_attention(d["q"], d["k"], d["v"], s["k"], shift, s["h"], s["w"])
        cases.check_packed(out, g[tag]["out"], TOL, tag)
        # block decomposition with exact fp64 attention per block
        b, c, h, w = s["b"], s["c"], s["h"], s["w"]
        sh, sw = ((h // s["k"]) // 2, (w // s["k"]) // 2) if shift else (0, 0)
        q4, k4, v4 = (d[n].double().view(b, h, w, c) for n in ("q", "k", "v"))
        dec = torch.empty_like(q4)
        for (r0, r1) in _axis_groups(h, s["k"], sh):
            for (c0, c1) in _axis_groups(w, s["k"], sw):
                qb, kb, vb = (t[:, r0:r1, c0:c1].reshape(b, -1, c) for t in (q4, k4, v4))
                p = torch.softmax(qb @ kb.transpose(1, 2) / c ** 0.5, -1)
                dec[:, r0:r1, c0:c1] = (p @ vb).view(b, r1 - r0, c1 - c0, c)
        cases.check_packed(dec.view(b, h * w, c).float(), g[tag]["out"], TOL, tag + " blocks")


@pytest.mark.parametrize("name", list(cases.F2B_CASES))
def test_f2b_transformer_layer(golden, name):
    """The oracle restatement of TransformerLayer.forward (self-attention and cross-attention + FFN blocks, plain and
    shifted windows) against vectors recorded from the reference, with the gradients of the token rows."""
    g = golden(name)
    s = cases.F2B_CASES[name]
    d = cases.f2b_inputs(s)
    for no_ffn in (True, False):
        for shift in (False, True):
            tag = ("self" if no_ffn else "cross") + ("_shift" if shift else "_plain")
            src, tgt = d["source"].clone().requires_grad_(True), d["target"].clone().requires_grad_(True)
            out = O.transformer_layer(src, tgt, d["params"], no_ffn, s["k"], shift, s["h"], s["w"])
            cases.check_packed(out, g[tag]["out"], TOL, tag)
            (out * d["wout"]).sum().backward()
            cases.check_packed(src.grad, g[tag]["dsource"], GTOL, tag + " dsource")
            cases.check_packed(tgt.grad, g[tag]["dtarget"], GTOL, tag + " dtarget")


# ---- the chained path (model.py:92-97 + gmflow.py:81-162) against hook captures from the unmodified reference CoUpdater
CHAIN_TOL = 2e-4       # twelve layers of fp32 re-association in front of near-one-hot softmaxes


def _check_chain(out, g):
    errs = {}
    for k in cases.CHAIN_KEYS:
        errs[k] = cases.check_packed(out[k], g[k], CHAIN_TOL, k)
    return errs


def test_chain_randn(golden):
    s = cases.CHAIN_CASES["chain_randn"]
    d = cases.chain_inputs(s)
    out = O.motion_chain(d["gm"], d["seg"], cases.chain_params(s["pseed"]), want=("ab", "feat", "flow_pred", "flow_prop", "mask", "corr1"))
    _check_chain(out, golden("chain_randn"))


def test_chain_insitu_c1(golden):
    g = golden("chain_insitu")
    out = O.motion_chain(g["gm"], g["seg"], cases.chain_params(), want=("ab", "feat", "flow_pred", "flow_prop", "mask", "corr1"))
    _check_chain(out, g)


# ---- training side (config c5): unFlowLoss.compute_loss and the training-mode chain against the unmodified reference
@pytest.mark.parametrize("name", list(cases.LOSS_CASES))
def test_unflow_loss(golden, name):
    g = golden(name)
    d = cases.loss_inputs(cases.LOSS_CASES[name])
    flows = [f.clone().requires_grad_(True) for f in d["flows"]]
    loss = O.unflow_loss(flows, d["images"])
    assert abs(float(loss) - g["loss"]) <= 1e-5 * abs(g["loss"])
    loss.backward()
    for f, gd in zip(flows, g["dflows"]):
        cases.check_packed(f.grad, gd, 1e-4, "dflow")


def test_train_chain(golden):
    g = golden("trainchain")
    s = cases.TRAIN_CHAIN_CASE
    d = cases.train_chain_inputs(s)
    P = {k: v.clone().requires_grad_(not k.startswith("GMFlow") and "running" not in k) for k, v in cases.chain_params(s["pseed"]).items()}
    gm, seg = d["gm"].clone().requires_grad_(True), d["seg"].clone().requires_grad_(True)
    ffw, fbw, corr, fea_new = O.motion_chain_train(gm, seg, P)
    assert len(ffw) == g["n_flows"] == 2
    for f, gf in zip(ffw, g["flow_fw"]):
        cases.check_packed(f, gf, 2e-4, "flow_fw")
    cases.check_packed(fea_new, g["fea_new"], 2e-4, "fea_new")
    lflow = O.unflow_loss([torch.cat((ffw[i], fbw[i]), 1) for i in range(2)], d["images"])
    loss = lflow + (fea_new * d["wseg"]).sum()
    assert abs(float(lflow) - g["loss_flow"]) <= 2e-4 * abs(g["loss_flow"])
    loss.backward()
    # the feeder-side gradients carry the fp32 noise of the matching path (cases.TRAIN_CHAIN_CASE): 2 % at this scale
    cases.check_packed(gm.grad, g["dgm"], 8e-2, "dgm")
    cases.check_packed(seg.grad, g["dseg"], 2e-2, "dseg")
    for k in cases.TRAIN_GRAD_KEYS:
        if k == "conv_corr.0.bias":                      # analytically zero: a BatchNorm on batch statistics follows
            assert P[k].grad.abs().max() < 1e-3
            continue
        cases.check_packed(P[k].grad, g["dparams"][k], 8e-2 if k.startswith(cases.TRAIN_NOISY) else 5e-3, k)
