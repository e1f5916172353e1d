#!/usr/bin/env python
"""bench.py -- frame-pairs/s of the EMIP motion-stream hot path on B200 (BASELINE.json metric, config c3).

Workload (BASELINE.json configs[2], "c3"): EMIP_short inference, **64 synthetic frame pairs at 352x352 in total**,
batch-sharded over the N GPUs (64 / 32 / 16 / 8 pairs per GPU: strong scaling).  One "step" = one pass of the CHAINED hot
path over the global batch -- everything ``CoUpdater.forward`` does between the backbones and the decoder (reference
model/EMIP_short/model.py:92-97 + motion/gmflow/gmflow.py:81-162, eval):

    camouflaged feeder x2 -> position add -> FeatureTransformer (6 self + 6 cross/FFN layers) -> global matching ->
    flow propagation -> upsampler convs + convex x8 upsampling -> conv_corr (re-associated, BN + ReLU folded, 3x3) ->
    motion collector

on post-backbone features (the PVT-v2 / CNN-encoder backbones are out of scope), through
``emip_b200.chain.MotionChain`` = our CUDA kernels behind the C ABI, replayed as one CUDA graph per step.

    python bench.py [--gpus N] [--steps K] [--warmup W]          our arm
    python bench.py --impl reference ...                          CPU arm (the oracle's chain on host cores, bounded sample)

N > 1 is launched by torchrun (one rank per GPU; no data-path collective: inference shards by frame pairs -- the only
collectives are the timing barrier and the max-over-ranks).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

GLOBAL_PAIRS = 64
C, H, W = 128, 44, 44
N = H * W
GM_STD, SEG_STD = 2.2, 1.0          # feature scales measured inside the model
N_INPUT_SETS = 4                    # rotated: inputs + the step's intermediates exceed the 126 MB L2 at every shard size
CPU_SAMPLE_PAIRS = 2                # bounded sample of the c3 batch for the CPU arms
# K1 sub-record (BASELINE.json configs[1], "c2")
NCU_MLP_DRAM_BYTES, NCU_MLP_ROWS = 0.4874e9, 247808    # measured once per kernel change with ncu (profiles/r4v_mlp_fused_full.txt)
K1_B = 16
K1_ALG_FLOP_PER_PAIR = 2.0 * N * N * C + 8.0 * N * N          # S counted once
K1_EXEC_MMA_FLOP_PER_PAIR = 2.0 * N * N * C * 3 * 2           # bf16 hi/lo split (x3), both directions (x2)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the roofline sub-records and baselines (scaling runs)")
    ap.add_argument("--no-graph", action="store_true", help="launch the chain kernel by kernel instead of replaying a CUDA graph")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def workload_config(n_gpus):
    per = GLOBAL_PAIRS // n_gpus
    return {"workload": "c3: EMIP_short inference, chained motion-stream hot path (feeder x2 -> FeatureTransformer -> global "
                        "matching -> flow propagation -> convex upsampling -> conv_corr -> motion collector) on post-backbone "
                        "features of 64 synthetic 352x352 frame pairs, batch-sharded",
            "global_pairs": GLOBAL_PAIRS, "pairs_per_gpu": per, "tokens": N, "channels": C,
            "l2_policy": f"inputs rotate over {N_INPUT_SETS} sets; inputs + per-step intermediates exceed the 126 MB L2",
            "arithmetic": "fp32 in / fp32 out; every contraction = three bf16 tcgen05 MMAs (hi.hi + lo.hi + hi.lo) accumulated in "
                          "fp32 (the 'bf16 operand storage, hi/lo split' mode of SURVEY.md 8d: meets the fp32 1e-3 tolerance); "
                          "fp32 softmax / LayerNorm",
            "sharding": f"frame pairs over {n_gpus} GPU(s), no data-path collective", "weights": "random-init (reference initialisers, seed 123)"}


# ---------------------------------------------------------------------------------------------------------- CPU arms
def cpu_chain_arm(n_runs, pairs):
    """The reference's algorithm for the chain restated on the CPU (oracle port), all host threads, bounded sample."""
    import torch
    import cases
    from oracle import restate as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = cases.chain_params()
    gm = cases.randn(400, (2 * pairs, C, H, W), GM_STD)
    seg = cases.randn(401, (2 * pairs, C, H, W), SEG_STD)
    times = []
    with torch.no_grad():
        O.motion_chain(gm, seg, P)                          # warm-up
        for _ in range(n_runs):
            t0 = time.perf_counter()
            O.motion_chain(gm, seg, P)
            times.append(time.perf_counter() - t0)
    return times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, cores = cpu_chain_arm(max(1, min(args.steps, 5)), CPU_SAMPLE_PAIRS)
    ms = 1e3 * sum(times) / len(times)
    v = CPU_SAMPLE_PAIRS / (ms / 1e3)
    sample = (f"{len(times)} passes of the chain over {CPU_SAMPLE_PAIRS} of the {GLOBAL_PAIRS} pairs (oracle port, torch CPU fp32, "
              f"{cores} threads); pairs/s does not depend on the sample size (per-pair work is independent)")
    emit({
        "impl": "reference", "metric": "frame-pairs/s", "value": v, "unit": "frame-pairs/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "frame-pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_JSON_FD = None


def claim_stdout():
    """Keep fd 1 for the ONE JSON line: everything else that writes to stdout from here on -- including C-level prints such as
    NCCL's version banner under torchrun -- goes to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(1 if _JSON_FD is None else _JSON_FD, data)


def bind_numa(local_rank):
    """Pin this rank's host threads to the cores next to its GPU (pinned staging buffers are then first-touched there)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [i * 64 + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
        if cpus and world > 1:            # deal the GPU-local cores to the ranks that share them
            share = [c for i, c in enumerate(cpus) if i % world == local_rank % world] or cpus
            os.sched_setaffinity(0, share)
            return len(share)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from emip_b200 import _lib
    from emip_b200._lib import I, SZ, ptr
    from emip_b200._ws import workspace
    from emip_b200.chain import MotionChain, GraphedChain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    host_cores = bind_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if GLOBAL_PAIRS % world:
        raise SystemExit(f"--gpus {world} does not divide the global batch of {GLOBAL_PAIRS} pairs")
    per = GLOBAL_PAIRS // world
    L = _lib.lib()
    _lib.check(L.emip_device_check(), "emip_device_check")

    K, Wm = args.steps, max(3, args.warmup)
    torch.manual_seed(123)                                             # configs/configs.yaml:69
    chain = MotionChain().to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    sets = [(GM_STD * torch.randn(2 * per, C, H, W, device=dev, generator=g), SEG_STD * torch.randn(2 * per, C, H, W, device=dev, generator=g))
            for _ in range(N_INPUT_SETS)]
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        chain(*sets[0])                                                # first call: function attributes, prepared weights
        torch.cuda.synchronize()
        if args.no_graph:
            graphs = None
            step = lambda i: chain(*sets[i % N_INPUT_SETS])
            launches = None
        else:
            graphs = [GraphedChain(chain, gm, seg, pool=None) for gm, seg in sets[:1]]
            graphs += [GraphedChain(chain, gm, seg, pool=graphs[0].pool()) for gm, seg in sets[1:]]
            step = lambda i: graphs[i % N_INPUT_SETS].replay()
            launches = graphs[0].kernel_nodes

        # ---------------- value: inputs resident in HBM, one chain pass per step ----------------
        for i in range(Wm):
            step(i)
        barrier()
        with ClockSampler(local) as clk:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(K):
                step(i)
            e1.record(stream)
            barrier()
        ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / K
    value = GLOBAL_PAIRS / (ms_step / 1e3)

    # ---------------- e2e: host buffers; H2D of the step's features + chain + D2H of its results inside the timed region ----
    # Triple-buffered: the H2D of step i+1 and the D2H of step i-1 (two copy streams) overlap the kernels of step i; the host reads
    # every step's results, two steps behind the launches, before that step's buffers are reused.
    nb = 3                                                              # buffers in flight: H2D(i+1) | kernels(i) | D2H(i-1), host reads step i-2
    pin = lambda t_: torch.empty(t_.shape, dtype=t_.dtype).pin_memory().copy_(t_)
    h_in = [(pin(sets[j][0].cpu()), pin(sets[j][1].cpu())) for j in range(nb)]
    with torch.no_grad():
        if graphs is not None:
            eg = graphs[:nb]
            d_in = [(gr.gm, gr.seg) for gr in eg]
            run_e = lambda j: eg[j].replay()
            outs = [gr.outputs for gr in eg]
        else:
            d_in = [(torch.empty_like(sets[0][0]), torch.empty_like(sets[0][1])) for _ in range(nb)]
            outs = [None] * nb

            def run_e(j):
                outs[j] = chain(*d_in[j])
        ref_out = graphs[0].outputs if graphs is not None else chain(*sets[0])
    res_names = ("flow_fw", "flow_bw", "fea_new")
    res_idx = (0, 1, 3)
    h_out = [[torch.empty(ref_out[i].shape, dtype=torch.float32).pin_memory() for i in res_idx] for _ in range(nb)]
    h2d_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    h2d_done = [torch.cuda.Event() for _ in range(nb)]
    comp_done = [torch.cuda.Event() for _ in range(nb)]
    d2h_done = [torch.cuda.Event() for _ in range(nb)]
    Ke = max(4, min(K, 16))

    def issue_h2d(i):
        j = i % nb
        h2d_stream.wait_event(comp_done[j])           # the kernels that last read these input buffers have finished
        with torch.cuda.stream(h2d_stream):
            d_in[j][0].copy_(h_in[j][0], non_blocking=True)
            d_in[j][1].copy_(h_in[j][1], non_blocking=True)
            h2d_done[j].record(h2d_stream)

    def e2e_loop(n):
        issue_h2d(0)
        for i in range(n):
            j = i % nb
            if i + 1 < n:
                issue_h2d(i + 1)
            stream.wait_event(h2d_done[j])
            stream.wait_event(d2h_done[j])            # the previous results of this buffer have left the device
            run_e(j)
            comp_done[j].record(stream)
            d2h_stream.wait_event(comp_done[j])
            with torch.cuda.stream(d2h_stream):
                for k_, idx in enumerate(res_idx):
                    h_out[j][k_].copy_(outs[j][idx], non_blocking=True)
                d2h_done[j].record(d2h_stream)
            if i >= nb - 1:
                d2h_done[(i - (nb - 1)) % nb].synchronize()  # the caller reads step i-2's results on the host
        for k_ in range(max(0, n - (nb - 1)), n):
            d2h_done[k_ % nb].synchronize()

    with torch.no_grad():
        for ev in comp_done + d2h_done:
            ev.record(stream)
        torch.cuda.synchronize()
        e2e_loop(3)
        reps = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            e2e_loop(Ke)
            torch.cuda.synchronize()
            reps.append(time.perf_counter() - t0)
        e2e_s = sorted(reps)[1]
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = GLOBAL_PAIRS * Ke / te.item()
    h2d = world * sum(t_.numel() * 4 for t_ in h_in[0])
    d2h = world * sum(t_.numel() * 4 for t_ in h_out[0])

    line = {
        "metric": "frame-pairs/s", "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(world), "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": "frame-pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": Ke, "passes": 3, "stat": "median of the passes", "results_read": list(res_names),
                "host_gbs": (h2d + d2h) * Ke / te.item() / 1e9, "host_cores_bound": host_cores},
        "gpu_launches": (launches * K * 1 if launches is not None else None),
        "launches_per_step": launches, "cuda_graph": graphs is not None,
    }

    if rank == 0 and not args.no_extras:
        pk = peaks()
        import chain_time
        with torch.no_grad():
            st = chain_time.stage_times(chain, sets[0][0], sets[0][1], iters=5)
        line["chain_stage_ms"] = {k: round(v, 4) for k, v in st.items()}
        line["roofline"] = dominant_kernel_roofline(torch, chain, per, dev, pk)
        line["roofline_feature_transformer"] = ft_roofline(torch, chain, per, dev, pk)
        line["roofline_k1_c2"] = k1_roofline(torch, dev, pk, max(K, 50))
        import k3_bench
        line["flow_warp_roofline"] = {"kernel": "flow_warp_staged_kernel<fwd|bwd> (TMA-staged; smooth flow) / flow_warp_border3_kernel (direct gather; noisy flow)",
                                      "bound": "hbm", "unit": "GB/s", "peak": pk["hbm"],
                                      "workload": "B=64, 3x352x352 fp32 (254 MB fwd / 317 MB bwd > L2)", "flows": k3_bench.run(dev, pk["hbm"])}
        # ---------------- the reference's op sequence as eager torch on this GPU (cuBLAS / cuDNN / ATen, fp32, TF32 off) ----
        try:
            import eager_chain
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            P = {k: v for k, v in chain.state_dict().items()}
            eb = min(per, 16)
            idx = torch.cat((torch.arange(eb), per + torch.arange(eb))).to(dev)
            gm, seg = sets[0][0][idx].contiguous(), sets[0][1][idx].contiguous()
            with torch.no_grad():
                for _ in range(2):
                    eager_chain.chain(gm, seg, P)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    eager_chain.chain(gm, seg, P)
                e1.record()
                torch.cuda.synchronize()
            ems = e0.elapsed_time(e1) / 3
            line["gpu_eager_baseline"] = {"value": eb / ems * 1e3, "unit": "frame-pairs/s", "ms_per_step": ems, "pairs": eb,
                                          "what": "the reference's op sequence for the same chain as eager torch on this GPU "
                                                  "(tools/eager_chain.py: cuBLAS / cuDNN / ATen, fp32, TF32 off)"}
        except Exception as e:  # noqa: BLE001 -- a baseline must not take the measurement down
            line["gpu_eager_baseline"] = {"error": repr(e)[:200]}
        # ---------------- CPU baseline (oracle port of the chain) on this box's host cores ----------------
        if not args.no_cpu_baseline and world == 1:
            try:
                os.sched_setaffinity(0, range(os.cpu_count()))
            except Exception:
                pass
            times, cores = cpu_chain_arm(3, CPU_SAMPLE_PAIRS)
            cv = CPU_SAMPLE_PAIRS / (sum(times) / len(times))
            line["cpu_baseline"] = {"value": cv, "unit": "frame-pairs/s", "cores": cores, "kind": "port",
                                    "sample": f"{len(times)} passes of the chain over {CPU_SAMPLE_PAIRS} of the {GLOBAL_PAIRS} pairs, torch CPU fp32"}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def dominant_kernel_roofline(torch, chain, per, dev, pk):
    """The chain's dominant kernel is the FeatureTransformer's feed-forward network (256 -> 1024 -> GELU -> 128 + LayerNorm +
    residual, transformer.py:139-147, :175-180; mlp_fused_kernel, six launches per pass, a quarter of the step) on the shard's
    2 * pairs * 1936 token rows: the public call (operand split of the fp32 rows + the fused kernel) timed alone with CUDA events
    on its stream, algorithmic FLOPs / time."""
    from emip_b200.transformer_layer import mlp_tm
    lay = chain.GMFlow.transformer.layers[0].cross_attn_ffn
    Lr = 2 * per * N
    xs = [torch.randn(Lr, 256, device=dev) for _ in range(3)]
    res = torch.randn(Lr, 128, device=dev)
    call = lambda i: mlp_tm(xs[i % 3], lay.mlp[0].weight, lay.mlp[2].weight, lay.norm2.weight, lay.norm2.bias, lay.norm2.eps, residual=res)
    with torch.no_grad():
        for i in range(3):
            call(i)
        torch.cuda.synchronize()
        n = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            call(i)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    alg = 2.0 * Lr * (256 * 1024 + 1024 * 128)
    ach = alg / (ms * 1e-3) / 1e12
    return {"kernel": "mlp_fused_kernel (FeatureTransformer feed-forward call: operand split of the fp32 rows + ONE fused kernel, hidden rows kept in TMEM)",
            "bound": "tensor", "achieved": ach,
            "peak": pk["tf"], "unit": "TFLOP/s", "frac": ach / pk["tf"],
            "traffic": NCU_MLP_DRAM_BYTES * Lr / NCU_MLP_ROWS, "traffic_unit": "bytes per launch of the fused kernel",
            "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of mlp_fused_kernel at 247808 rows "
                              "(profiles/r4v_mlp_fused_full.txt: 0.382 GB read + 0.105 GB written, 546 us, tensor pipe 52 % of elapsed), scaled "
                              "by rows; algorithmic: 0.25 (rows, bf16 hi | lo) + 0.13 (residual) + 0.13 (out) GB -- the two-launch version "
                              "it replaces moved 2.59 GB (hidden rows out and back in); the operand split in front of the kernel in THIS call "
                              "(0.25 GB read + 0.25 GB written) does not exist inside the FeatureTransformer call, where the rows arrive pre-split",
            "peak_source": pk["src"] + " burst bf16 (cuBLAS)", "call_ms": ms, "rows": Lr,
            "executed_mma_tflops": 3 * ach, "executed_mma_frac": 3 * ach / pk["tf"],
            "note": "achieved = algorithmic 2*L*(256*1024 + 1024*128) FLOP / time; executed = x3 (bf16 hi/lo split keeps fp32 accuracy)"}


def ft_roofline(torch, chain, per, dev, pk):
    """The FeatureTransformer call (emip_feature_transformer_fwd: 36 gemm_tc_kernel + 6 mlp_fused_kernel + 12 attn_fwd_tc_kernel launches)
    = half of the step: algorithmic FLOPs of its twelve layers / CUDA-event time of the call."""
    from emip_b200.chain import feature_transformer_tokens
    maps = 2 * per
    Lr = maps * N
    xs = [torch.randn(maps, N, C, device=dev) for _ in range(2)]
    with torch.no_grad():
        for i in range(2):
            feature_transformer_tokens(xs[i], chain.GMFlow.transformer, H, W, 2)
        torch.cuda.synchronize()
        n = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            feature_transformer_tokens(xs[i % 2], chain.GMFlow.transformer, H, W, 2)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    lin = 2.0 * Lr * (8 * 128 * 128 + 256 * 1024 + 1024 * 128)                    # per block: q k v merge x2, mlp
    pairs_plain = 4 * 484 * 484                                                  # token pairs attended per map, unshifted layer
    pairs_shift = 484 * 484 + 4 * 242 * 242 + 4 * 121 * 121                      # shifted layer (block decomposition of the mask)
    att = lambda pr: 2 * (2.0 * pr * 128 * 2) * maps                              # two attention layers per block, QK^T and PV
    alg = 3 * (lin + att(pairs_plain)) + 3 * (lin + att(pairs_shift))
    ach = alg / (ms * 1e-3) / 1e12
    return {"call": "emip_feature_transformer_fwd (6 blocks)", "bound": "tensor", "achieved": ach, "peak": pk["tf"], "unit": "TFLOP/s",
            "frac": ach / pk["tf"], "call_ms": ms, "rows": Lr, "executed_mma_tflops": 3 * ach, "executed_mma_frac": 3 * ach / pk["tf"],
            "note": "algorithmic FLOPs of the six block pairs (linear layers + window attention); executed = x3 (bf16 hi/lo split)"}


def k1_roofline(torch, dev, pk, K):
    """BASELINE.json configs[1] (c2): the fused global-matching kernel alone at 16 pairs, with `corr` emitted."""
    from emip_b200 import _lib
    from emip_b200._lib import I, SZ, ptr
    from emip_b200._ws import workspace
    L = _lib.lib()
    g = torch.Generator(device=dev).manual_seed(99)
    n_sets = 8
    sets = [(4.1 * torch.randn(K1_B, C, H, W, device=dev, generator=g), 4.1 * torch.randn(K1_B, C, H, W, device=dev, generator=g))
            for _ in range(n_sets)]
    stream = torch.cuda.current_stream()
    L.emip_global_matching_workspace.restype = ctypes.c_size_t
    nbytes = L.emip_global_matching_workspace(I(K1_B), I(C), I(H), I(W))
    wss = [workspace(nbytes, dev) for _ in range(n_sets)]
    flow = torch.empty((2 * K1_B, 2, H, W), device=dev)
    corr = torch.empty((K1_B, N, H, W), device=dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    def kernel_only(i, flags, corr_t):
        f0, f1 = sets[i % n_sets]
        _, wp, wn = wss[i % n_sets]
        _lib.check(L.emip_global_matching_fwd(ptr(f0), ptr(f1), ptr(flow), ptr(corr_t), None, ctypes.c_void_p(wp), SZ(wn), I(K1_B), I(C),
                                              I(H), I(W), I(1), I(flags), sp), "fwd")
    for i in range(n_sets):
        kernel_only(i, 0, corr)            # fills every workspace with its operand split
    res = {}
    for name, corr_t, fl in (("with_corr", corr, 2), ("flow_only", None, 2), ("bf16_with_corr", corr, 6), ("bf16_flow_only", None, 6)):
        for i in range(5):
            kernel_only(i, fl, corr_t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(K):
            kernel_only(i, fl, corr_t)
        e1.record(stream)
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / K
    k_ms = res["with_corr"]
    ach = K1_ALG_FLOP_PER_PAIR * K1_B / (k_ms * 1e-3) / 1e12
    ex = K1_EXEC_MMA_FLOP_PER_PAIR * K1_B / (k_ms * 1e-3) / 1e12
    return {"kernel": "match_tc_fwd_kernel", "workload": "c2: 16 pairs, 44x44 tokens, C=128, bidirectional, corr emitted", "bound": "tensor",
            "achieved": ach, "peak": pk["tf"], "unit": "TFLOP/s", "frac": ach / pk["tf"], "launch_ms": k_ms,
            "launch_ms_flow_only": res["flow_only"], "launch_ms_bf16_mode": res["bf16_with_corr"],
            "launch_ms_bf16_mode_flow_only": res["bf16_flow_only"], "executed_mma_tflops": ex, "executed_mma_frac": ex / pk["tf"],
            "corr_write_gbs": K1_B * N * N * 4 / (k_ms * 1e-3) / 1e9, "pairs_per_s": K1_B / (k_ms * 1e-3),
            "note": "achieved counts S once (2BN^2C+8BN^2); executed = x3 (bf16 hi/lo split) x2 (both directions)"}


if __name__ == "__main__":
    main()
