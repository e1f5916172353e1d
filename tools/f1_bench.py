#!/usr/bin/env python
"""f1 micro-benchmark: conv_corr[0] on the never-materialised cost volume (two tcgen05 GEMMs) vs F.conv2d on the
materialised corr tensor (cuDNN, fp32 and TF32), model shape: 44x44 tokens, C=128, O=968."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from emip_b200.conv_corr import conv_corr_first_layer
from emip_b200.matching import global_correlation_softmax


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    H = W = 44; C = 128; O = 968; N = H * W
    res = {}
    w = torch.randn(O, N, 3, 3, device="cuda") * (9 * N) ** -0.5
    b = torch.randn(O, device="cuda") * 0.1
    for B in (1, 16):
        f0 = 4.1 * torch.randn(B, C, H, W, device="cuda")
        f1 = 4.1 * torch.randn(B, C, H, W, device="cuda")
        with torch.no_grad():
            t_ours = timeit(lambda: conv_corr_first_layer(f0, f1, w, b))
            corr = global_correlation_softmax(f0, f1, True)[2]
            torch.backends.cudnn.allow_tf32 = False
            t_fp32 = timeit(lambda: F.conv2d(corr, w, b, padding=1), iters=5, warm=2)
            torch.backends.cudnn.allow_tf32 = True
            t_tf32 = timeit(lambda: F.conv2d(corr, w, b, padding=1), iters=5, warm=2)
            t_a1_corr = timeit(lambda: global_correlation_softmax(f0, f1, True))
            t_a1_flow = timeit(lambda: global_correlation_softmax(f0, f1, True, return_corr=False))
        alg = 2.0 * B * (O * 9 * N * C + O * N * 9 * C)           # the two GEMMs
        ref_flops = 2.0 * B * O * N * 9 * N                         # the convolution it replaces
        res[f"B{B}"] = dict(ours_ms=t_ours, conv_fp32_ms=t_fp32, conv_tf32_ms=t_tf32, a1_with_corr_ms=t_a1_corr,
                            a1_flow_only_ms=t_a1_flow, gemm_gflop=alg / 1e9, executed_mma_tflops=3 * alg / t_ours / 1e9,
                            replaced_conv_gflop=ref_flops / 1e9)
        print(f"B={B:2d}: re-associated conv_corr[0] {t_ours*1e3:8.1f} us ({3*alg/t_ours/1e9:6.0f} TFLOP/s executed bf16 MMA, "
              f"{alg/1e9:.1f} GFLOP algorithmic)  | F.conv2d on corr: fp32 {t_fp32*1e3:8.1f} us, TF32 {t_tf32*1e3:8.1f} us "
              f"({ref_flops/1e9:.0f} GFLOP)  | a1 with corr {t_a1_corr*1e3:6.1f} us, flow only {t_a1_flow*1e3:6.1f} us")
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
