// Fused feed-forward network of the FeatureTransformer's cross-attention layer on the tensor cores; see mlp_fused.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

// y = res + LayerNorm_128(GELU(X W1^T) W2^T) * gamma + beta       (transformer.py:175-176, :180; mlp = Linear, GELU, Linear, no bias)
struct MlpFusedArgs {
  const void* x_hi; const void* x_lo; int ldx;      // bf16 rows [L][ldx]: the [source | message] operand, 256 columns used
  const void* w1;                                   // bf16 [hid][hi 256 | lo 256]  (gemm_tc_split_b of mlp.0.weight [hid, 256])
  const void* w2;                                   // bf16 [128][hi hid | lo hid]  (gemm_tc_split_b of mlp.2.weight [128, hid])
  int L, hid;                                       // rows; hidden width (multiple of 128)
  const float* gamma; const float* beta; float eps; // norm2
  const float* res; int ldr;                        // residual rows (may be NULL)
  float* y; int ldy;                                // fp32 output rows (may be NULL)
  void* out_hi; void* out_lo; int out_ld;           // the output rows as bf16 hi / lo (may be NULL): the next block's row operand
};

bool mlp_fused_supported(const MlpFusedArgs& a);
int mlp_fused_tc(const MlpFusedArgs& a, cudaStream_t st);
