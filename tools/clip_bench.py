#!/usr/bin/env python
"""Config c4: EMIP_long over 5-frame synthetic clips, clips sharded over the GPUs -- the hot-path part of
``Model_long.forward`` (reference model/EMIP_long/model_long.py:68-117; test_long.py:29-37 calls it frame by frame, index 0..4):

    per frame pair (batch 1 by signature):  the chained short-term path (MotionChain: feeder x2 -> FeatureTransformer -> matching
    -> flow propagation -> upsampling -> conv_corr -> collector)                                                [model_long.py:70-96]
    index >= 1:  historical-feature prompt = Memory read over the last T <= 5 frames' keys / values (a5, LTM.py:49-68),
    then the long-term motion collector injector1(fea_2, memory)                                             [model_long.py:98-113]

on post-backbone features; the key / value / long_dr convolutions and the decoder are out of scope (library code): keys and
values are synthetic tensors of the right shape, and the 256-channel memory read-out is cut to its first 128 channels where
the reference applies long_dr.  Frames inside a clip are sequential (memory_k / memory_v); a GPU runs ``--streams`` clips
concurrently, each as one CUDA graph per frame index.

    python tools/clip_bench.py [--clips 64] [--streams 4]            (torchrun for N GPUs: clips are dealt to the ranks)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


class ClipWorker:
    """One clip in flight: static buffers + one captured graph per frame index (T = 0 .. 4 memory entries)."""

    def __init__(self, chain, inj_long, mem, dev, stream, pool):
        from emip_b200.chain import _count_kernel_nodes
        self.stream = stream
        g = torch.Generator(device=dev).manual_seed(11)
        r = lambda *s, sc=1.0: sc * torch.randn(*s, device=dev, generator=g)
        self.gm, self.seg = r(2, 128, 44, 44, sc=2.2), r(2, 128, 44, 44)
        self.keys, self.vals = r(1, 128, 5, 44, 44, sc=1.5), r(1, 128, 5, 44, 44)      # FIFO of the last frames' keys / values
        self.qk, self.qv = r(1, 128, 44, 44, sc=1.5), r(1, 128, 1, 44, 44)
        self.graphs, self.nodes = [], []
        with torch.no_grad(), torch.cuda.stream(stream):
            for T in range(5):
                def frame(T=T):
                    out = chain(self.gm, self.seg)                                       # short-term path, batch 1
                    if T == 0:
                        return out
                    m, _ = mem(self.keys[:, :, :T], self.vals[:, :, :T], self.qk, self.qv)      # [1, 256, 44, 44]
                    return out, inj_long(self.seg[1:2], m[:, :128].contiguous())        # long_dr (conv, out of scope) stands here
                for _ in range(2):
                    frame()
                stream.synchronize()
                gr = torch.cuda.CUDAGraph(keep_graph=True)
                with torch.cuda.graph(gr, pool=pool, stream=stream):
                    self.out = frame()
                self.nodes.append(_count_kernel_nodes(gr))
                gr.replay()
                self.graphs.append(gr)
                pool = pool or gr.pool()
        self.pool = pool
        stream.synchronize()

    def run_clip(self):
        with torch.cuda.stream(self.stream):
            for T in range(5):
                self.graphs[T].replay()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--json", default=None)
    ap.add_argument("--pdl", action="store_true", help="capture the graphs with programmatic dependent launch (A / B: it costs 28 %% "
                                                        "with four concurrent clip streams, profiles/r5q_clips_pdl_ab.txt)")
    args = ap.parse_args()
    if args.pdl:
        from emip_b200 import _lib
        _lib.lib().emip_set_programmatic_launch(1)
    from emip_b200.chain import MotionChain
    from emip_b200.injector import Injector
    from emip_b200.memory import Memory
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    per = args.clips // world
    torch.manual_seed(123)
    chain = MotionChain().to(dev).eval()
    inj_long = Injector().to(dev).eval()
    mem = Memory()
    res = {}
    for ns in sorted({1, args.streams}):
        workers, pool = [], None
        for i in range(ns):
            w = ClipWorker(chain, inj_long, mem, dev, torch.cuda.Stream(device=dev), None)
            workers.append(w)
        torch.cuda.synchronize()

        def run_all():
            for c in range(0, per, ns):
                for w in workers[: min(ns, per - c)]:
                    w.run_clip()
        run_all()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main_s = torch.cuda.current_stream()
        e0.record(main_s)
        for w in workers:
            w.stream.wait_event(e0)
        for _ in range(args.reps):
            run_all()
        for w in workers:
            main_s.wait_stream(w.stream)
        e1.record(main_s)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[ns] = dict(ms_per_pass=ms.item(), clips_per_s=args.clips / ms.item() * 1e3, frames_per_s=5 * args.clips / ms.item() * 1e3,
                       ms_per_frame_per_stream=ms.item() * ns / (5 * per), kernel_nodes_per_frame=workers[0].nodes)
        del workers
    if rank == 0:
        out = {"workload": "c4: EMIP_long hot path over 5-frame synthetic clips (short-term chain at batch 1 + memory read + long-term "
                           "collector per frame), clips sharded over the GPUs", "n_gpus": world, "clips": args.clips, "clips_per_gpu": per,
               "by_concurrent_clips_per_gpu": res}
        print(json.dumps(out))
        if args.json:
            with open(args.json, "w") as f:
                json.dump(out, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
