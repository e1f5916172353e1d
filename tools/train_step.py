#!/usr/bin/env python
"""Config c5: one training step of the chained path per iteration -- forward (training mode), unFlowLoss + a fixed cotangent on
the motion collector's output (decoder + hybrid_e_loss are out of scope), backward, gradient all-reduce of the trainable set over
NCCL overlapped with the backward (BucketedGradAllReduce), elementwise gradient clamp (utils.py:1-11), AdamW step.

    python tools/train_step.py [--pairs 64] [--steps 10]                      1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step.py ...   N GPUs (global batch sharded)

Prints one JSON line on rank 0: frame-pairs/s, ms per step, all-reduce ms / exposed ms / bus GB/s, and (N > 1) the parity of the
all-reduced gradients against the same shards processed one after the other on rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=64, help="global batch (frame pairs)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bucket-mb", type=float, default=25.0)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from emip_b200.chain import MotionChain
    from emip_b200.dist import BucketedGradAllReduce
    from emip_b200.flow_loss import unflow_loss
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    per = args.pairs // world
    H = W = 44
    torch.manual_seed(123)
    m = MotionChain().to(dev).train().freeze_like_reference()
    params = [p for p in m.parameters() if p.requires_grad]
    red = BucketedGradAllReduce(params, bucket_mb=args.bucket_mb)
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=1e-7)               # configs.yaml:62-66

    def shard(r):
        g = torch.Generator(device=dev).manual_seed(777 + r)
        return dict(gm=2.2 * torch.randn(2 * per, 128, H, W, device=dev, generator=g), seg=torch.randn(2 * per, 128, H, W, device=dev, generator=g),
                    images=0.5 * torch.randn(per, 6, 8 * H, 8 * W, device=dev, generator=g),
                    wseg=0.01 * torch.randn(per, 128, H, W, device=dev, generator=g))

    def fwd_bwd(d):
        ffw, fbw, corr, fea_new = m.forward_train(d["gm"], d["seg"])
        lflow = unflow_loss([torch.cat((ffw[i], fbw[i]), 1) for i in range(2)], d["images"])[0]        # train.py:53-58
        loss = lflow + (fea_new * d["wseg"]).sum()
        loss.backward()
        return loss

    data = shard(rank)
    parity = None
    if world > 1:
        # ---- NCCL parity: all-reduced gradients == mean over the shards processed one after the other on one GPU
        red.zero_grad()
        fwd_bwd(data)
        red.finish()
        torch.cuda.synchronize()
        got = red.flat.clone()
        if rank == 0:
            for h in red._hooks:
                h.remove()
            ref = torch.zeros_like(got)
            for r in range(world):
                red.flat.zero_()
                fwd_bwd(shard(r))
                ref += red.flat / world
            parity = ((got - ref).norm() / ref.norm()).item()
            red._hooks = [p.register_post_accumulate_grad_hook(red._ready) for p in red.params]
        dist.barrier()

    def step():
        red.zero_grad()
        loss = fwd_bwd(data)
        red.finish()
        red.flat.clamp_(-0.5, 0.5)                                             # utils.py:1-11 clip_gradient(optimizer, 0.5)
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    comm = exposed = 0.0
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    comm, exposed = red.stats()                                                # of the last step
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # ---- the as-trained all-reduce size (reference: 100.42 M trainable floats incl. the PVT backbone), alone
    big_ms = None
    if world > 1:
        big = torch.zeros(100_420_000, device=dev)
        for _ in range(2):
            dist.all_reduce(big)
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(5):
            dist.all_reduce(big)
        b1.record()
        torch.cuda.synchronize()
        big_ms = b0.elapsed_time(b1) / 5
    if rank == 0:
        bus = lambda nbytes, t_ms: 2 * (world - 1) / world * nbytes / (t_ms * 1e-3) / 1e9 if world > 1 and t_ms else None
        out = {"workload": "c5: training step of the chained path (fwd + unFlowLoss + bwd + all-reduce + clamp + AdamW)", "n_gpus": world,
               "global_pairs": args.pairs, "pairs_per_gpu": per, "ms_per_step": ms.item(), "pairs_per_s": args.pairs / ms.item() * 1e3,
               "loss": float(loss), "trainable_params": red.numel, "allreduce_bytes": red.nbytes, "buckets": len(red.buckets),
               "allreduce_ms_last_step": comm, "allreduce_exposed_ms_last_step": exposed, "allreduce_bus_gbs": bus(red.nbytes, comm),
               "grad_parity_vs_sequential_shards": parity,
               "allreduce_401mb_ms": big_ms, "allreduce_401mb_bus_gbs": bus(401.68e6, big_ms)}
        print(json.dumps(out))
        if args.json:
            with open(args.json, "w") as f:
                json.dump(out, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
