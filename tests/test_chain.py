"""The chained hot path (emip_b200.chain.MotionChain = model.py:92-97 + gmflow.py:81-162 on our kernels) against hook
captures from the unmodified reference CoUpdater (tests/golden/chain_*.pt) and against the CPU oracle."""
import pytest
import torch

import cases
from oracle import restate as O

FLOW_TOL = 1e-3        # north_star: flow within 1e-3 relative in fp32 (model features: measured 1.4e-4 ... 2.1e-4)
FEAT_TOL = 1e-3
# Unstructured (iid gaussian) features make the matching ill-conditioned: every FeatureTransformer block amplifies a
# perturbation of its input ~1.7x (attention logits up to +-93) and the two near-one-hot softmaxes of a1 / a2 another ~6x.
# The reference's own fp32 arithmetic is 8e-5 ... 9e-5 away from fp64 on these flows (tools/chain_err.py; test_oracle_golden);
# the split-bf16 GEMMs (2^-17 operands instead of 2^-24) land at 7e-4 ... 1.2e-3.  The bound for this synthetic-feature
# case is therefore 2e-3; the model-feature case (config c1 in situ) keeps the 1e-3 of north_star.
FLOW_TOL_RANDN = 2e-3


def _chain(P, **kw):
    from emip_b200.chain import MotionChain
    m = MotionChain(**kw)
    missing, unexpected = m.load_state_dict(P, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    return m


def test_chain_state_dict_keys_are_the_references():
    """CPU: the chain owns exactly the path's parameters under CoUpdater's state_dict keys."""
    from emip_b200.chain import MotionChain
    m = MotionChain()
    keys = {k for k in m.state_dict() if not k.endswith("num_batches_tracked")}
    P = cases.chain_params()
    assert keys == set(P), (sorted(keys - set(P))[:5], sorted(set(P) - keys)[:5])
    for k, v in m.state_dict().items():
        if k in P:
            assert tuple(v.shape) == tuple(P[k].shape), k
    from oracle import ref_shim
    if ref_shim.available():
        ref_shim.install()
        from model.EMIP_short.model import CoUpdater
        ref = CoUpdater(ref_shim.model_args()).state_dict()
        assert keys <= set(ref)
        assert all(tuple(ref[k].shape) == tuple(P[k].shape) for k in keys)


def test_chain_refuses_cpu_tensors():
    m = _chain(cases.chain_params(hw=16 * 16, corr_mid=64), hw=16 * 16, corr_mid=64).eval()
    from emip_b200._lib import EmipError
    with torch.no_grad(), pytest.raises(EmipError):
        m(torch.zeros(2, 128, 16, 16), torch.zeros(2, 128, 16, 16))


def test_token_row_entry_points_refuse_cpu_and_wrong_operands():
    """The hand-over wrappers fail loudly instead of falling back: CPU tensors, fp32 rows where the bf16 hi | lo operand is expected."""
    from emip_b200 import chain as ch
    from emip_b200._lib import EmipError
    ft = ch._FeatureTransformer()
    with pytest.raises(EmipError):
        ch.feature_transformer_tokens(torch.zeros(2, 64, 128), ft, 8, 8, 2, want_split=True)
    ffa = _chain(cases.chain_params(hw=16 * 16, corr_mid=64), hw=16 * 16, corr_mid=64).GMFlow.feature_flow_attn
    with pytest.raises((EmipError, TypeError)):
        ch.flow_attention_tokens(torch.zeros(2, 64, 256, dtype=torch.bfloat16), ffa, torch.zeros(2, 2, 64))
    if torch.cuda.is_available():
        with pytest.raises(TypeError):                           # fp32 rows are not the pre-split operand
            ch.flow_attention_tokens(torch.zeros(2, 64, 256, device="cuda"), ffa.cuda(), torch.zeros(2, 2, 64, device="cuda"))


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _run_and_check(m, gm, seg, g, tol_flow=FLOW_TOL, tol_feat=FEAT_TOL):
    B2, C, H, W = gm.shape
    with torch.no_grad():
        ffw, fbw, corr, fea_new, loc = m(gm.cuda(), seg.cuda(), want=("ab", "feat_tok", "flow_pred", "flow_prop", "mask"))
    feat = loc["feat_tok"].view(B2, H, W, C).permute(0, 3, 1, 2)
    big = 1e9
    errs = {
        "ab": cases.check_packed(loc["ab"], g["ab"], big, "ab"),
        "feat": cases.check_packed(feat, g["feat"], big, "feat"),
        "flow_pred": cases.check_packed(loc["flow_pred"], g["flow_pred"], big, "flow_pred"),
        "flow_prop": cases.check_packed(loc["flow_prop"], g["flow_prop"], big, "flow_prop"),
        "mask": cases.check_packed(loc["mask"], g["mask"], big, "upsampler mask"),
        "flow_fw": cases.check_packed(ffw, g["flow_fw"], big, "flow_fw"),
        "flow_bw": cases.check_packed(fbw, g["flow_bw"], big, "flow_bw"),
        "corr": cases.check_packed(corr, g["corr"], big, "conv_corr output"),
        "fea_new": cases.check_packed(fea_new, g["fea_new"], big, "injector1 output"),
    }
    print("chain rel-L2 vs reference:", {k: f"{v:.2e}" for k, v in errs.items()})
    tol = dict(ab=tol_feat, feat=tol_feat, corr=tol_feat, fea_new=tol_feat)
    bad = {k: v for k, v in errs.items() if not v <= tol.get(k, tol_flow)}
    assert not bad, f"over tolerance: {bad} (all: {errs})"
    return errs, loc


@pytest.mark.gpu
def test_chain_randn_vs_reference(golden):
    s = cases.CHAIN_CASES["chain_randn"]
    d = cases.chain_inputs(s)
    m = _chain(cases.chain_params(s["pseed"])).cuda().eval()
    _run_and_check(m, d["gm"], d["seg"], golden("chain_randn"), tol_flow=FLOW_TOL_RANDN)


@pytest.mark.gpu
def test_chain_insitu_c1_vs_reference(golden):
    """config c1: the features the real backbones produce for one seeded 352x352 pair (captured by hooks inside CoUpdater)."""
    g = golden("chain_insitu")
    m = _chain(cases.chain_params()).cuda().eval()
    _run_and_check(m, g["gm"], g["seg"], g)


@pytest.mark.gpu
def test_chain_conv_corr_first_layer_token_major(golden):
    """conv_corr[0] alone (no folded BatchNorm) from the chain's token rows against the reference's conv_corr[0] output."""
    import ctypes
    from emip_b200 import _lib
    from emip_b200._lib import I, SZ, ptr, stream_ptr
    from emip_b200._ws import workspace
    from emip_b200.conv_corr import _prepared_weight
    g = golden("chain_randn")
    s = cases.CHAIN_CASES["chain_randn"]
    d = cases.chain_inputs(s)
    P = cases.chain_params(s["pseed"])
    m = _chain(P).cuda().eval()
    with torch.no_grad():
        *_, loc = m(d["gm"].cuda(), d["seg"].cuda(), want=("feat_tok",))
    tok = loc["feat_tok"]
    B, H, W, C, Oc = s["b"], s["h"], s["w"], 128, 968
    L = _lib.lib()
    L.emip_conv_corr_workspace.restype = ctypes.c_size_t
    ws, wp, wn = workspace(L.emip_conv_corr_workspace(I(B), I(C), I(H), I(W), I(Oc)), tok.device)
    out = torch.empty((B, Oc, H, W), device=tok.device)
    wprep = _prepared_weight(m.conv_corr[0].weight)[1]
    _lib.check(L.emip_conv_corr_fwd_ex(ptr(tok[:B]), ptr(tok[B:]), ctypes.c_void_p(wprep), ptr(m.conv_corr[0].bias.detach()), None, None,
                                       I(0), I(0), ptr(out), ctypes.c_void_p(wp), SZ(wn), I(B), I(C), I(H), I(W), I(Oc), stream_ptr()),
               "emip_conv_corr_fwd_ex")
    cases.check_packed(out, g["corr1"], FEAT_TOL, "conv_corr[0] output")


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,b", [(16, 16, 1), (12, 20, 2)])
def test_chain_small_vs_oracle(h, w, b):
    """Fresh seeded inputs at sizes the oracle finishes in seconds (windows of 8x8 / 6x10 tokens, ragged pixel tiles)."""
    P = cases.chain_params(seed=9, hw=h * w, corr_mid=72)
    gm = cases.randn(301, (2 * b, 128, h, w), 2.2)
    seg = cases.randn(302, (2 * b, 128, h, w), 1.0)
    ref = O.motion_chain(gm, seg, P, want=("ab", "feat", "flow_pred", "flow_prop", "mask"))
    m = _chain(P, hw=h * w, corr_mid=72).cuda().eval()
    with torch.no_grad():
        ffw, fbw, corr, fea_new, loc = m(gm.cuda(), seg.cuda(), want=("ab", "feat_tok", "flow_pred", "flow_prop", "mask"))
    feat = loc["feat_tok"].view(2 * b, h, w, 128).permute(0, 3, 1, 2)
    errs = dict(ab=_rel(loc["ab"], ref["ab"]), feat=_rel(feat, ref["feat"]), flow_pred=_rel(loc["flow_pred"], ref["flow_pred"]),
                flow_prop=_rel(loc["flow_prop"], ref["flow_prop"]), mask=_rel(loc["mask"], ref["mask"]), flow_fw=_rel(ffw, ref["flow_fw"]),
                flow_bw=_rel(fbw, ref["flow_bw"]), corr=_rel(corr, ref["corr"]), fea_new=_rel(fea_new, ref["fea_new"]))
    print("chain small rel-L2 vs oracle:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert all(v <= FLOW_TOL_RANDN for v in errs.values()), errs
    assert all(errs[k] <= 1e-3 for k in ("ab", "feat", "corr", "fea_new")), errs


@pytest.mark.gpu
def test_chain_batch_consistency():
    """B = 3 pairs in one call == the same pairs one at a time (kernels pick different schedules / tile counts)."""
    P = cases.chain_params()
    m = _chain(P).cuda().eval()
    gm = cases.randn(311, (6, 128, 44, 44), 2.2).cuda()
    seg = cases.randn(312, (6, 128, 44, 44), 1.0).cuda()
    with torch.no_grad():
        full = m(gm, seg)
        for i in range(3):
            idx = torch.tensor([i, 3 + i], device="cuda")
            one = m(gm[idx].contiguous(), seg[idx].contiguous())
            for a, b_ in zip(one, full):
                assert _rel(a, b_[i:i + 1]) < 2e-4


@pytest.mark.gpu
@pytest.mark.parametrize("b,h,w,k", [(1, 44, 44, 2), (2, 16, 24, 2), (1, 12, 18, 3)])
def test_feature_transformer_fused_call_vs_oracle_and_per_layer_path(b, h, w, k):
    """The one-call FeatureTransformer (window-ordered operands written by the projection GEMMs, pre-split merge / MLP
    operands, no concat1 / cat) against the fp64 oracle and against the per-layer path built from the separately tested ops."""
    from emip_b200 import chain as ch
    P = cases.chain_params(seed=5)
    m = ch._FeatureTransformer()
    m.load_state_dict(O.sub_params(P, "GMFlow.transformer."))
    m = m.cuda()
    x = cases.randn(321, (2 * b, h * w, 128), 2.0)
    c0 = x.double()
    c1 = torch.cat(c0.chunk(2, 0)[::-1], 0)
    P64 = {kk: v.double() for kk, v in P.items()}
    for i in range(6):
        ps, pc = (O.sub_params(P64, f"GMFlow.transformer.layers.{i}.{n}.") for n in ("self_attn", "cross_attn_ffn"))
        c0 = O.transformer_layer(c0, c0, ps, True, k, i % 2 == 1, h, w)
        c0 = O.transformer_layer(c0, c1, pc, False, k, i % 2 == 1, h, w)
        c1 = torch.cat(c0.chunk(2, 0)[::-1], 0)
    with torch.no_grad():
        xg = x.cuda()
        fused = ch.feature_transformer_tokens(xg, m, h, w, k)
        per = xg
        for blk in m.layers:
            per = ch.transformer_block(blk, per, h, w, k)
    e_f, e_p, e_fp = _rel(fused, c0), _rel(per, c0), _rel(fused, per)
    print(f"feature transformer {b}x{h}x{w} k={k}: fused vs fp64 {e_f:.2e}, per-layer vs fp64 {e_p:.2e}, fused vs per-layer {e_fp:.2e}")
    assert e_f < 3e-4 and e_p < 3e-4 and e_fp < 3e-4


@pytest.mark.gpu
def test_train_chain_vs_reference(golden):
    """Config c5's training step on the chained path: forward in training mode (two flow predictions, BatchNorm on batch
    statistics), unFlowLoss on our kernels, backward through every op's own backward kernel -- against the same pass through the
    unmodified reference modules (tests/golden/trainchain.pt).  Gradients that pass through the matching path are compared at
    the fp32 noise floor of that path (see cases.TRAIN_CHAIN_CASE), the others at 5e-3."""
    from emip_b200.flow_loss import unflow_loss
    g = golden("trainchain")
    s = cases.TRAIN_CHAIN_CASE
    d = cases.train_chain_inputs(s)
    m = _chain(cases.chain_params(s["pseed"])).cuda().train().freeze_like_reference()
    gm, seg = d["gm"].cuda().requires_grad_(True), d["seg"].cuda().requires_grad_(True)
    ffw, fbw, corr, fea_new = m.forward_train(gm, seg)
    assert len(ffw) == len(fbw) == g["n_flows"] == 2
    errs = {f"flow_fw[{i}]": cases.check_packed(ffw[i], g["flow_fw"][i], 2e-3, f"flow_fw[{i}]") for i in range(2)}
    errs["fea_new"] = cases.check_packed(fea_new, g["fea_new"], 1e-3, "fea_new")
    lflow = unflow_loss([torch.cat((ffw[i], fbw[i]), 1) for i in range(2)], d["images"].cuda())[0]
    loss = lflow + (fea_new * d["wseg"].cuda()).sum()
    assert abs(float(lflow) - g["loss_flow"]) <= 2e-3 * abs(g["loss_flow"]), (float(lflow), g["loss_flow"])
    loss.backward()
    errs["dgm"] = cases.check_packed(gm.grad, g["dgm"], 1e9, "dgm")
    errs["dseg"] = cases.check_packed(seg.grad, g["dseg"], 1e9, "dseg")
    grads = dict(m.named_parameters())
    tol = {"dgm": 0.2, "dseg": 0.1}
    for k in cases.TRAIN_GRAD_KEYS:
        if k == "conv_corr.0.bias":
            assert grads[k].grad.abs().max() < 1e-2
            continue
        errs[k] = cases.check_packed(grads[k].grad, g["dparams"][k], 1e9, k)
        # the two temperatures are a 2-element tensor (one sub-sampled number): its error is pure noise of the matching path
        tol[k] = 0.5 if k.endswith("temperature") else 0.2 if k.startswith(cases.TRAIN_NOISY) else 5e-3
    print("train chain rel-L2 vs reference:", {k: f"{v:.2e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if k in tol and not v <= tol[k]}
    assert not bad, bad
    assert all(p.grad is None for n, p in m.named_parameters() if n.startswith("GMFlow"))


def test_fuse_chain_glue_reproduces_the_reference_forward_on_cpu():
    """Row (g), the part that needs the reference: ``dropin.fuse_chain`` replaces model.py:92-97 inside the UNMODIFIED reference
    ``CoUpdater`` (its own backbones, dr1-3 and decoder).  With the chain object's kernel call swapped for the CPU oracle's
    ``motion_chain`` (the GPU tests pin our kernels to that oracle), the fused model must return the reference's own mask logits
    and flows: same stacking of the two frames, same outputs wired into dr1 / the decoder, argmax mask bit-exact."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not mounted")
    ref_shim.install()
    from model.EMIP_short.model import CoUpdater
    from emip_b200 import dropin
    torch.manual_seed(123)
    net = CoUpdater(ref_shim.model_args()).eval()
    net.load_state_dict(cases.chain_params(), strict=False)
    im1, im2 = cases.randn(0, (1, 3, 352, 352)), cases.randn(1, (1, 3, 352, 352))
    with torch.no_grad():
        mask_ref, ffw_ref, fbw_ref = net(im1, im2)
    dropin.fuse_chain(net)
    P = {k: v.detach() for k, v in net.state_dict().items()}

    def oracle_forward(gm, seg, want=()):
        out = O.motion_chain(gm, seg, P)
        return out["flow_fw"], out["flow_bw"], out["corr"], out["fea_new"]
    net._emip_chain.forward = oracle_forward
    with torch.no_grad():
        mask, ffw, fbw = net(im1, im2)
    assert len(ffw) == len(ffw_ref) == 1 and len(fbw) == 1
    assert _rel(ffw[0], ffw_ref[0]) < 2e-4 and _rel(fbw[0], fbw_ref[0]) < 2e-4
    assert _rel(mask, mask_ref) < 1e-4
    assert torch.equal(mask > 0, mask_ref > 0)                                     # argmax mask (1-channel logit > 0), SURVEY F5
    # in training mode the original forward runs
    net.train()
    assert net.forward.__wrapped__ is not None


@pytest.mark.gpu
@pytest.mark.parametrize("b,c0,c1,o,h,w,lay0,lay1,relu,scale", [
    (2, 2, 128, 256, 44, 44, 1, 0, True, False),      # upsampler[0]: cat(flow [CN], tokens [NC]) -> 256, ReLU
    (1, 968, 0, 128, 44, 44, 1, 1, False, False),     # conv_corr[3]
    (3, 37, 0, 72, 12, 20, 0, 0, False, True),        # ragged: Cin / O not multiples of the tile sizes, token-major input, BN-style affine
    (2, 70, 11, 130, 16, 16, 1, 1, True, True),       # two channel-major sources, partial M tile
])
def test_conv3x3_vs_library_convolution(b, c0, c1, o, h, w, lay0, lay1, relu, scale):
    """emip_conv3x3_fwd (shifted-TMA-box im2col, split-bf16 tcgen05) against the fp64 convolution of the concatenated input."""
    from emip_b200 import chain as ch
    x0 = cases.randn(401, (b, c0, h, w), 1.5)
    x1 = cases.randn(402, (b, c1, h, w), 0.7) if c1 else None
    wt = cases.randn(403, (o, c0 + c1, 3, 3), (9 * (c0 + c1)) ** -0.5)
    sh = cases.randn(404, (o,), 0.3)
    sc = (1 + 0.2 * cases.randn(405, (o,))) if scale else None
    xin = x0 if x1 is None else torch.cat((x0, x1), 1)
    ref = torch.nn.functional.conv2d(xin.double(), wt.double(), None, padding=1)
    ref = ref * (sc.double().view(1, -1, 1, 1) if scale else 1.0) + sh.double().view(1, -1, 1, 1)
    if relu:
        ref = ref.clamp_min(0)
    tok = lambda t: t.flatten(2).transpose(1, 2).contiguous()
    g0 = (x0 if lay0 == 1 else tok(x0)).cuda()
    g1 = None if x1 is None else (x1 if lay1 == 1 else tok(x1)).cuda()
    wp = ch.prepare_conv3x3(wt.cuda())
    out = ch.conv3x3(g0, lay0, g1, lay1, wp, o, h, w, scale=None if sc is None else sc.cuda(), shift=sh.cuda(), relu=relu)
    assert _rel(out, ref) < 5e-5           # K = 9 * 968 products per output at the largest case: 3e-5


@pytest.mark.gpu
@pytest.mark.parametrize("b,h,w", [(1, 44, 44), (5, 16, 24), (19, 44, 44)])
def test_fused_mlp_kernel_vs_two_gemm_launches(b, h, w):
    """mlp_fused.cu (hidden rows kept in TMEM, LayerNorm in its epilogue) against the two gemm_tc launches it replaces, through
    the whole FeatureTransformer call (six feed-forward networks deep): same products, another summation order in the second GEMM."""
    from emip_b200 import chain as ch, _lib
    P = cases.chain_params(seed=11)
    x = cases.randn(341, (2 * b, h * w, 128), 2.0).cuda()
    L = _lib.lib()
    # one block: the kernels' own difference (summation order); six blocks: the same, amplified by the (random-weight) network --
    # both paths sit at the same distance from the fp64 oracle (test_feature_transformer_fused_call_vs_oracle_and_per_layer_path)
    for n_blocks, tol in ((1, 2e-6), (6, 3e-4)):
        m = ch._FeatureTransformer()
        m.load_state_dict(O.sub_params(P, "GMFlow.transformer."))
        m.layers = m.layers[:n_blocks]
        m = m.cuda()
        with torch.no_grad():
            try:
                L.emip_debug_gemm_wide_tiles(4)                      # bit 2: two-launch feed-forward network
                two, two_split = ch.feature_transformer_tokens(x, m, h, w, 2, want_split=True)
            finally:
                L.emip_debug_gemm_wide_tiles(0)
            fused, fused_split = ch.feature_transformer_tokens(x, m, h, w, 2, want_split=True)
        err = _rel(fused, two)
        print(f"fused mlp vs two launches, {n_blocks} block(s), {2 * b} maps of {h}x{w}: rel-L2 {err:.2e}, max |d| {(fused - two).abs().max().item():.2e}")
        assert err < tol
        hi = fused.to(torch.bfloat16)
        assert torch.equal(fused_split[..., :128], hi) and torch.equal(fused_split[..., 128:], (fused - hi.float()).to(torch.bfloat16))


@pytest.mark.gpu
@pytest.mark.parametrize("b,h,w", [(1, 44, 44), (3, 16, 24)])
def test_presplit_handover_of_the_transformer_rows_is_bit_identical(b, h, w):
    """The bf16 hi | lo rows the last LayerNorm epilogue writes (emip_feature_transformer_fwd_ex out_split) are the split of the
    fp32 output, and a1 (EMIP_FLAG_PRESPLIT) / a2 (emip_flow_attn_tokens_fwd) fed with them return exactly what they return
    from the fp32 rows through their own split passes."""
    from emip_b200 import chain as ch
    from emip_b200.flow_attn import flow_attention_core
    P = cases.chain_params(seed=9)
    m = ch._FeatureTransformer()
    m.load_state_dict(O.sub_params(P, "GMFlow.transformer."))
    m = m.cuda()
    ffa = _chain(P).GMFlow.feature_flow_attn.cuda()
    x = cases.randn(331, (2 * b, h * w, 128), 2.0).cuda()
    with torch.no_grad():
        out, split = ch.feature_transformer_tokens(x, m, h, w, 2, want_split=True)
        assert torch.equal(out, ch.feature_transformer_tokens(x, m, h, w, 2))
        hi = out.to(torch.bfloat16)
        lo = (out - hi.float()).to(torch.bfloat16)
        assert torch.equal(split[..., :128], hi) and torch.equal(split[..., 128:], lo)
        flow_a = ch.global_matching_tokens(out, b, h, w)
        flow_b = ch.global_matching_tokens(split, b, h, w)
        assert torch.equal(flow_a, flow_b)
        q = ch.linear_tm_bias(out, ffa.q_proj.weight, ffa.q_proj.bias)
        k = ch.linear_tm_bias(q, ffa.k_proj.weight, ffa.k_proj.bias)
        v = flow_a.view(2 * b, 2, h * w)
        prop_a = flow_attention_core(q, k, v)
        prop_b = ch.flow_attention_tokens(split, ffa, v)
        assert torch.equal(prop_a, prop_b)


@pytest.mark.gpu
def test_tokens_from_cn_and_bias_linear_and_kv_swap():
    """The small pieces of the chain against torch: transpose + position add, nn.Linear with bias on token rows, and the
    cross-attention call that reads the keys / values of the other batch half (no swapped copy)."""
    from emip_b200 import chain as ch
    x = cases.randn(411, (3, 128, 10, 14))
    pos = ch.window_position(10, 14, 2, 128, torch.device("cuda", 0))
    ref = (x + pos.cpu().view(1, 128, 10, 14)).flatten(2).transpose(1, 2)
    assert torch.equal(ch.tokens_from_cn(x.cuda(), pos).cpu(), ref)
    assert torch.equal(pos.cpu().view(128, 10, 14), O.position_embedding_sine(5, 7, 64).repeat(1, 2, 2))
    t = cases.randn(412, (2, 300, 128), 2.0)
    wt, bs = cases.randn(413, (128, 128), 128 ** -0.5), cases.randn(414, (128,), 0.2)
    assert _rel(ch.linear_tm_bias(t.cuda(), wt.cuda(), bs.cuda()), t.double() @ wt.double().T + bs.double()) < 1e-5
    q, k, v = (cases.randn(415 + i, (4, 16 * 24, 128), 1.5 if i < 2 else 1.0) for i in range(3))
    swap = lambda z: torch.cat(z.chunk(2, 0)[::-1], 0)
    for shift in (False, True):
        ref = O.split_window_attention(q.double(), swap(k).double(), swap(v).double(), 2, shift, 16, 24)
        out = ch.window_attention(q.cuda(), k.cuda(), v.cuda(), 2, shift, 16, 24, kv_swap_halves=True)
        assert _rel(out, ref) < 5e-5
