"""Generate golden vectors by running the UNMODIFIED reference (zhangxin06/EMIP).

Run in the build container, where /root/reference is mounted:

    python tests/golden/make_golden.py

Writes tests/golden/*.pt.  The GPU box never sees the reference, so these files
(plus the seeded input builders in cases.py) are what pins the oracle there.
Recorded with torch 2.11.0+cu128 on CPU, fp32, deterministic seeds.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle import ref_shim  # noqa: E402

ref_shim.install()
torch.set_num_threads(os.cpu_count())

from model.EMIP_short.motion.gmflow.matching import global_correlation_softmax  # noqa: E402
from model.EMIP_short.motion.gmflow.transformer import FeatureFlowAttention  # noqa: E402
from model.EMIP_short.motion.PromptInteract import Injector  # noqa: E402
from model.EMIP_long.LTM import Memory  # noqa: E402
from loss.warp_utils import flow_warp  # noqa: E402


def save(name, obj):
    path = os.path.join(HERE, name + ".pt")
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def gen_a1():
    for name, s in cases.A1_CASES.items():
        d = cases.a1_inputs(s)
        f0 = d["f0"].clone().requires_grad_(True)
        f1 = d["f1"].clone().requires_grad_(True)
        flow, prob, corr = global_correlation_softmax(f0, f1, True)
        loss = (flow * d["wflow"]).sum() + (corr * d["wcorr"]).sum()
        loss.backward()
        save(name, dict(spec=s, flow=cases.pack(flow, False), corr=cases.pack(corr, s["full"]),
                        prob_rowsum_err=float((prob.sum(-1) - 1).abs().max()),
                        df0=cases.pack(f0.grad, s["full"]), df1=cases.pack(f1.grad, s["full"])))


def gen_a2():
    for name, s in cases.A2_CASES.items():
        d = cases.a2_inputs(s)
        m = FeatureFlowAttention(s["c"])
        m.load_state_dict({k: d[k] for k in ("q_proj.weight", "q_proj.bias", "k_proj.weight", "k_proj.bias")})
        x = d["x"].clone().requires_grad_(True)
        out = m(x, d["flow"])
        (out * d["wout"]).sum().backward()
        save(name, dict(spec=s, out=cases.pack(out, False), dx=cases.pack(x.grad, s["full"]),
                        dqw=cases.pack(m.q_proj.weight.grad, False), dqb=cases.pack(m.q_proj.bias.grad, False),
                        dkw=cases.pack(m.k_proj.weight.grad, False)))  # dL/d(k bias) == 0 analytically


def gen_a3():
    for name, s in cases.A3_CASES.items():
        d = cases.a3_inputs(s)
        x = d["x"].clone().requires_grad_(True)
        fl = d["flow"].clone().requires_grad_(True)
        out = flow_warp(x, fl, pad=s["pad"])
        (out * d["wout"]).sum().backward()
        save(name, dict(spec=s, out=cases.pack(out, False), dflow=cases.pack(fl.grad, False),
                        dx=cases.pack(x.grad, False)))
    # the non-contiguous channel-slice call pattern of loss_flow.py:90-91
    s = dict(b=2, c=3, h=12, w=10, sigma=4.0, pad="border", seed=35)
    flow4 = cases.randn(s["seed"], (s["b"], 4, s["h"], s["w"]), s["sigma"])
    x = cases.randn(s["seed"] + 1, (s["b"], 3, s["h"], s["w"]))
    save("a3_slice", dict(spec=s, out_fw=cases.pack(flow_warp(x, flow4[:, :2]), False),
                          out_bw=cases.pack(flow_warp(x, flow4[:, 2:]), False)))


def gen_a4():
    for name, s in cases.A4_CASES.items():
        d = cases.a4_inputs(s)
        m = Injector()
        m.transformer.load_state_dict(d["params"])
        x = d["x"].clone().requires_grad_(True)
        x1 = d["x1"].clone().requires_grad_(True)
        out = m(x, x1)
        (out * d["wout"]).sum().backward()
        g = {k: cases.pack(p.grad, False) for k, p in m.transformer.named_parameters()}
        save(name, dict(spec=s, out=cases.pack(out, s["full"]), dx=cases.pack(x.grad, s["full"]),
                        dx1=cases.pack(x1.grad, s["full"]), dparams=g))


def gen_a5():
    for name, s in cases.A5_CASES.items():
        d = cases.a5_inputs(s)
        t = {k: d[k].clone().requires_grad_(True) for k in ("m_in", "m_out", "q_in", "q_out")}
        out, p = Memory()(t["m_in"], t["m_out"], t["q_in"], t["q_out"])
        (out * d["wout"]).sum().backward()
        save(name, dict(spec=s, out=cases.pack(out, s["full"]),
                        **{"d" + k: cases.pack(v.grad, s["full"]) for k, v in t.items()}))


def gen_f4():
    """Convex x8 upsampling: the reference method GMFlow.upsample_flow with the upsampler conv replaced by a stub
    that returns the seeded mask (the conv itself is out of scope)."""
    from model.EMIP_short.motion.gmflow.gmflow import GMFlow

    class Stub:
        upsample_factor = 8

    for name, s in cases.F4_CASES.items():
        d = cases.f4_inputs(s)
        flow = d["flow"].clone().requires_grad_(True)
        mask = d["mask"].clone().requires_grad_(True)
        stub = Stub()
        stub.upsampler = lambda concat, m=mask: m
        out = GMFlow.upsample_flow(stub, flow, torch.zeros(s["b"], 128, s["h"], s["w"]))
        (out * d["wout"]).sum().backward()
        save(name, dict(spec=s, out=cases.pack(out, s["full"]), dflow=cases.pack(flow.grad, False),
                        dmask=cases.pack(mask.grad, s["full"])))


def gen_f1():
    """conv_corr[0] (model.py:59) on the corr tensor the reference's matching function returns."""
    from model.EMIP_short.motion.gmflow.matching import global_correlation_softmax
    for name, s in cases.F1_CASES.items():
        d = cases.f1_inputs(s)
        f0 = d["f0"].clone().requires_grad_(True)
        f1 = d["f1"].clone().requires_grad_(True)
        conv = torch.nn.Conv2d(s["h"] * s["w"], s["o"], 3, 1, 1)
        with torch.no_grad():
            conv.weight.copy_(d["weight"])
            conv.bias.copy_(d["bias"])
        corr = global_correlation_softmax(f0, f1, True)[2]
        out = conv(corr)
        (out * d["wout"]).sum().backward()
        save(name, dict(spec=s, out=cases.pack(out.detach(), False), df0=cases.pack(f0.grad, False), df1=cases.pack(f1.grad, False),
                        dw=cases.pack(conv.weight.grad, False), db=cases.pack(conv.bias.grad, False)))


def gen_f2():
    """single_head_split_window_attention (transformer.py:46-105) with and without the half-window shift."""
    import model.EMIP_short.motion.gmflow.transformer as T
    for name, s in cases.F2_CASES.items():
        d = cases.f2_inputs(s)
        full = name.endswith("full")
        rec = dict(spec=s)
        for shift in (False, True):
            q, k, v = (d[n].clone().requires_grad_(True) for n in ("q", "k", "v"))
            mask = None
            if shift:
                wh, ww = s["h"] // s["k"], s["w"] // s["k"]
                mask = T.generate_shift_window_attn_mask((s["h"], s["w"]), wh, ww, wh // 2, ww // 2, device=torch.device("cpu"))
            out = T.single_head_split_window_attention(q, k, v, num_splits=s["k"], with_shift=shift, h=s["h"], w=s["w"],
                                                       attn_mask=mask)
            (out * d["wout"]).sum().backward()
            tag = "shift" if shift else "plain"
            rec[tag] = dict(out=cases.pack(out.detach(), full), dq=cases.pack(q.grad, True), dk=cases.pack(k.grad, True),
                            dv=cases.pack(v.grad, True))
        save(name, rec)


def gen_f2b():
    """TransformerLayer.forward (transformer.py:151-180): self-attention layer (no_ffn) and cross-attention + FFN layer, plain and
    shifted windows, with the gradients of the two token tensors."""
    import model.EMIP_short.motion.gmflow.transformer as T
    for name, s in cases.F2B_CASES.items():
        d = cases.f2b_inputs(s)
        full = name.endswith("full")
        rec = dict(spec=s)
        wh, ww = s["h"] // s["k"], s["w"] // s["k"]
        mask = T.generate_shift_window_attn_mask((s["h"], s["w"]), wh, ww, wh // 2, ww // 2, device=torch.device("cpu"))
        for no_ffn in (True, False):
            for shift in (False, True):
                m = T.TransformerLayer(d_model=s["c"], nhead=1, attention_type="swin", no_ffn=no_ffn, ffn_dim_expansion=4,
                                       with_shift=shift)
                m.load_state_dict({k: v for k, v in d["params"].items() if not (no_ffn and (k.startswith("mlp") or k.startswith("norm2")))},
                                  strict=False)
                src, tgt = d["source"].clone().requires_grad_(True), d["target"].clone().requires_grad_(True)
                out = m(src, tgt, height=s["h"], width=s["w"], shifted_window_attn_mask=mask, attn_num_splits=s["k"])
                (out * d["wout"]).sum().backward()
                tag = ("self" if no_ffn else "cross") + ("_shift" if shift else "_plain")
                rec[tag] = dict(out=cases.pack(out.detach(), full), dsource=cases.pack(src.grad, full), dtarget=cases.pack(tgt.grad, full))
        save(name, rec)


def gen_f3b():
    """unFlowLoss.loss_photomatric (loss_flow.py:35-49) with gradients to the reconstruction."""
    from loss.loss_flow import unFlowLoss
    crit = unFlowLoss()
    for name, s in cases.F3B_CASES.items():
        d = cases.f3b_inputs(s)
        rec = d["rec"].clone().requires_grad_(True)
        loss = crit.loss_photomatric(d["im"], rec, d["mask"])
        loss.backward()
        save(name, dict(spec=s, loss=float(loss), drec=cases.pack(rec.grad, False)))


def gen_f3():
    from loss.warp_utils import get_occu_mask_backward, get_corresponding_map, mesh_grid
    for name, s in cases.F3_CASES.items():
        fl = cases.f3_inputs(s)["flow4"][:, 2:]                     # the slice loss_flow.py:95 passes
        base = mesh_grid(s["b"], s["h"], s["w"]).type_as(fl)
        save(name, dict(spec=s, mask=cases.pack(get_occu_mask_backward(fl, th=0.2), False),
                        corr_map=cases.pack(get_corresponding_map(base + fl), False)))


def gen_c1():
    """config c1: CoUpdater eval forward on one seeded 352x352 pair; capture the a1/a2 call in situ."""
    from model.EMIP_short.model import CoUpdater
    import model.EMIP_short.motion.gmflow.gmflow as gm
    torch.manual_seed(123)  # configs/configs.yaml:69
    net = CoUpdater(ref_shim.model_args()).eval()
    im1 = cases.randn(0, (1, 3, 352, 352))
    im2 = cases.randn(1, (1, 3, 352, 352))
    cap = {}
    orig = gm.global_correlation_softmax

    def spy(f0, f1, bidir):
        cap["f0"], cap["f1"] = f0.detach().clone(), f1.detach().clone()
        out = orig(f0, f1, bidir)
        cap["flow"], cap["corr"] = out[0].detach().clone(), out[2].detach().clone()
        return out

    def ffa_hook(mod, args, kwargs, out):
        cap["ffa_in_flow"] = args[1].detach().clone()
        cap["ffa_out"] = out.detach().clone()

    gm.global_correlation_softmax = spy
    h = net.GMFlow.feature_flow_attn.register_forward_hook(ffa_hook, with_kwargs=True)
    with torch.no_grad():
        mask, ffw, fbw = net(im1, im2)
    gm.global_correlation_softmax = orig
    h.remove()
    ffa = net.GMFlow.feature_flow_attn
    s_max = (cap["corr"].max().item(), cap["corr"].min().item(), cap["corr"].std().item())
    save("c1_insitu", dict(
        f0=cap["f0"], f1=cap["f1"], flow=cases.pack(cap["flow"], False), corr=cases.pack(cap["corr"], True),
        ffa_params={k: v.detach().clone() for k, v in ffa.state_dict().items()},
        ffa_out=cases.pack(cap["ffa_out"], False),
        mask=cases.pack(mask, True), flow_fw=cases.pack(ffw[-1], True), flow_bw=cases.pack(fbw[-1], True),
        stats=dict(corr_max_min_std=s_max, flow_abs_mean=cap["flow"].abs().mean().item(),
                   mask_mean=mask.mean().item(), mask_std=mask.std().item())))
    print("c1 stats", s_max, cap["flow"].abs().mean().item())


def _chain_model():
    from model.EMIP_short.model import CoUpdater
    torch.manual_seed(123)  # configs/configs.yaml:69
    net = CoUpdater(ref_shim.model_args()).eval()
    P = cases.chain_params()
    missing, unexpected = net.load_state_dict(P, strict=False)
    assert not unexpected, unexpected
    return net, P


def _chain_capture(net, run):
    """Runs ``run()`` with hooks that record the intermediates of the chained path inside the unmodified reference."""
    import model.EMIP_short.motion.gmflow.gmflow as gm
    cap = {"ab": []}
    orig = gm.global_correlation_softmax

    def spy(f0, f1, bidir):
        out = orig(f0, f1, bidir)
        cap["feat"] = torch.cat((f0, f1), 0).detach().clone()
        cap["flow_pred"] = out[0].detach().clone()
        return out

    hooks = [
        net.injector.register_forward_hook(lambda m, a, o: cap["ab"].append(o.detach().clone())),
        net.GMFlow.feature_flow_attn.register_forward_hook(lambda m, a, o: cap.__setitem__("flow_prop", o.detach().clone())),
        net.GMFlow.upsampler.register_forward_hook(lambda m, a, o: cap.__setitem__("mask", o.detach().clone())),
        net.conv_corr[0].register_forward_hook(lambda m, a, o: cap.__setitem__("corr1", o.detach().clone())),
        net.conv_corr.register_forward_hook(lambda m, a, o: cap.__setitem__("corr", o.detach().clone())),
        net.injector1.register_forward_hook(lambda m, a, o: cap.__setitem__("fea_new", o.detach().clone())),
    ]
    gm.global_correlation_softmax = spy
    try:
        with torch.no_grad():
            res = run()
    finally:
        gm.global_correlation_softmax = orig
        for h in hooks:
            h.remove()
    cap["ab"] = torch.cat(cap["ab"], 0)
    return cap, res


def gen_chain():
    """The chained hot path (model.py:92-97 + gmflow.py:81-162, eval) inside the unmodified reference CoUpdater whose
    path weights were overwritten with cases.chain_params(): (i) seeded post-backbone features, (ii) config c1 in situ --
    backbone features recorded by hooks from a full ``CoUpdater.forward`` on one seeded 352x352 pair."""
    net, P = _chain_model()
    for name, s in cases.CHAIN_CASES.items():
        d = cases.chain_inputs(s)
        B = s["b"]

        def run(d=d, B=B):
            a = net.injector(d["gm"][:B], d["seg"][:B])                    # model.py:92
            b = net.injector(d["gm"][B:], d["seg"][B:])                    # model.py:93
            flow_fw, flow_bw, corr = net.GMFlow([a], [b])                  # model.py:94
            corr = net.conv_corr(corr)                                     # model.py:96
            net.injector1(d["seg"][:B], corr)                              # model.py:97
            return flow_fw, flow_bw
        cap, (ffw, fbw) = _chain_capture(net, run)
        cap["flow_fw"], cap["flow_bw"] = ffw[-1], fbw[-1]
        save(name, dict(spec=s, **{k: cases.pack(cap[k], True) for k in cases.CHAIN_KEYS}))
    # in situ (config c1): the features the real backbones produce
    im1 = cases.randn(0, (1, 3, 352, 352))
    im2 = cases.randn(1, (1, 3, 352, 352))
    feats = {"gm": [], "seg": []}
    hb = [net.GMFlow.backbone.register_forward_hook(lambda m, a, o: feats["gm"].append(o[0].detach().clone())),
          net.backbone.feat_net.register_forward_hook(lambda m, a, o: feats["seg"].append(o[0].detach().clone()))]
    cap, (mask, ffw, fbw) = _chain_capture(net, lambda: net(im1, im2))
    for h in hb:
        h.remove()
    cap["flow_fw"], cap["flow_bw"] = ffw[-1], fbw[-1]
    save("chain_insitu", dict(gm=torch.cat(feats["gm"], 0), seg=torch.cat(feats["seg"], 0), seg_mask=cases.pack(mask, True),
                              **{k: cases.pack(cap[k], True) for k in cases.CHAIN_KEYS}))
    print("chain_insitu stats: feat std", cap["feat"].std().item(), "flow_pred |mean|", cap["flow_pred"].abs().mean().item(),
          "corr std", cap["corr"].std().item(), "fea_new std", cap["fea_new"].std().item(), "mask std", mask.std().item())


def gen_loss():
    """unFlowLoss.compute_loss (loss_flow.py:60-138) on seeded smooth flows: total loss and its gradients w.r.t. both flow entries."""
    from loss.loss_flow import unFlowLoss
    crit = unFlowLoss()
    for name, s in cases.LOSS_CASES.items():
        d = cases.loss_inputs(s)
        flows = [f.clone().requires_grad_(True) for f in d["flows"]]
        total, warp, smooth, mean_abs = crit.compute_loss(flows, d["images"])
        total.backward()
        save(name, dict(spec=s, loss=float(total), dflows=[cases.pack(f.grad, False) for f in flows]))


def gen_trainchain():
    """One training-mode pass of the chained path inside the unmodified reference modules (GMFlow.train(): two flow
    predictions; conv_corr's BatchNorm on batch statistics) + unFlowLoss + a fixed cotangent on the motion collector's output
    (standing in for decoder + hybrid_e_loss, out of scope), backward: loss and gradients of the trainable path parameters."""
    from loss.loss_flow import unFlowLoss
    net, P = _chain_model()
    s = cases.TRAIN_CHAIN_CASE
    d = cases.train_chain_inputs(s)
    B = s["b"]
    net.train()
    for n, p in net.named_parameters():
        p.requires_grad_(not ("GMFlow" in n))                         # train.py:340-342 (the unused dwconv / adaptor params aside)
    gm, seg = d["gm"].clone().requires_grad_(True), d["seg"].clone().requires_grad_(True)
    a = net.injector(gm[:B], seg[:B])
    b = net.injector(gm[B:], seg[B:])
    flow_fw, flow_bw, corr = net.GMFlow([a], [b])
    corr = net.conv_corr(corr)
    fea_new = net.injector1(seg[:B], corr)
    pairs = [torch.cat((flow_fw[i], flow_bw[i]), 1) for i in range(len(flow_fw))]                # train.py:55-57
    lflow = unFlowLoss().compute_loss(pairs, d["images"])[0]
    loss = lflow + (fea_new * d["wseg"]).sum()
    loss.backward()
    grads = dict(net.named_parameters())
    save("trainchain", dict(spec=s, loss_flow=float(lflow), loss=float(loss), n_flows=len(flow_fw),
                            flow_fw=[cases.pack(f, True) for f in flow_fw], fea_new=cases.pack(fea_new, True),
                            dgm=cases.pack(gm.grad, True), dseg=cases.pack(seg.grad, True),
                            dparams={k: cases.pack(grads[k].grad, True) for k in cases.TRAIN_GRAD_KEYS}))
    print("trainchain: loss_flow", float(lflow), "loss", float(loss))


if __name__ == "__main__":
    which = sys.argv[1:] or ["a1", "a2", "a3", "a4", "a5", "f1", "f2", "f2b", "f3", "f3b", "f4", "c1", "chain", "loss", "trainchain"]
    for w in which:
        globals()["gen_" + w]()
