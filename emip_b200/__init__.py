"""emip_b200 -- B200-native (sm_100a) implementation of EMIP's motion-stream hot path.

Host side: thin PyTorch wrappers (device memory, streams, autograd plumbing)
that mirror the reference's own function / nn.Module interfaces for the path:

    matching.global_correlation_softmax   <- model/EMIP_short/motion/gmflow/matching.py:8
    flow_attn.FeatureFlowAttention         <- model/EMIP_short/motion/gmflow/transformer.py:485
    warp.flow_warp                         <- loss/warp_utils.py:83
    injector.Injector                      <- model/EMIP_short/motion/PromptInteract.py:452
    memory.Memory                          <- model/EMIP_long/LTM.py:44
    window_attn.single_head_split_window_attention / single_head_full_attention   <- .../gmflow/transformer.py:46, :8
    transformer_layer.transformer_layer_forward   <- .../gmflow/transformer.py:151 (TransformerLayer.forward)
    conv_corr.CorrConv2d, upsample.upsample_flow, photometric.loss_photomatric, warp.get_occu_mask_backward   (the "next" rows)

All arithmetic runs in hand-written CUDA behind the C ABI of
include/emip_b200.h (emip_b200/_C/libemip_b200.so).  There is no CPU fallback.
"""
__version__ = "0.1.0"
