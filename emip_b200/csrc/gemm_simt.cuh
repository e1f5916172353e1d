// Small exact-fp32 batched GEMM building blocks used by the prompt-fusion (Injector) kernels.
// The injector's inputs/outputs must stay fp32-accurate, its GEMMs are
// 128..680-wide per-sample 1x1 convolutions on 1936 pixels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

// Y[b][m][n] = sum_k W(b)[m][k] * X'(b)[k][n]  (+ R[b][m][n])
//   W(b)[m][k] = w[b*w_stride_b + (w_trans ? k*ldw + m : m*ldw + k)]
//   X'(b)[k][n] = x[b*x_stride_b + k*ldx + n], optionally layer-normalised on load:
//                 (x - mean[b][n]) * rstd[b][n] * gamma[k] + beta[k]
struct GemmNN {
  const float* w; long long w_stride_b; int ldw; int w_trans;
  const float* x; long long x_stride_b; int ldx;
  const float* mean; const float* rstd; const float* gamma; const float* beta;   // all NULL = no LN; stats are [B][N]
  const float* res; long long res_stride_b; int ldr;                              // NULL = no residual
  float* y; long long y_stride_b; int ldy;
  int B, M, K, N;
  int accumulate;                                                                  // y += instead of y =
  const float* bias;                                                               // NULL or [M]: added per output row (tensor-core path only)
  int x_presplit;                                                                  // tensor-core path: the caller already wrote the bf16 hi | lo
                                                                                   // activation operand at gemm_nn_tc_act_operand(scratch, ...)
};
int gemm_nn(const GemmNN& a, cudaStream_t st);

// C[b][m][k] = sum_n A(b)[m][n] * B'(b)[k][n]     (contraction over the contiguous pixel axis)
//   A(b)[m][n] = a[b*a_stride_b + m*lda + n];  B'(b)[k][n] = bm[b*b_stride_b + k*ldb + n] with optional LN as above
struct GemmNT {
  const float* a; long long a_stride_b; int lda;
  const float* bm; long long b_stride_b; int ldb;
  const float* mean; const float* rstd; const float* gamma; const float* beta;
  float* c; long long c_stride_b; int ldc;
  int B, M, K, N;
  int a_act;                                                                       // 1: exact GELU applied to A on load (tensor-core path only)
  const float* a_aux;                                                              // a_act 2: A * GELU'(a_aux) on load, a_aux laid out like a
  // tensor-core path only, B == 1: A already split (bf16 hi, lo rows of pitch kpad(N)); C written as bf16 hi | lo rows of
  // pitch ldc_split (c_act 1: exact GELU first) instead of fp32 c -- the MLP of the FeatureTransformer blocks
  const void* a_hi_pre; const void* a_lo_pre;
  void* c_hi; void* c_lo; int ldc_split; int c_act;
  // tensor-core path only, K == 128: c = c_res + LayerNorm over the K outputs of a row (gamma, beta [K]; c_res rows ldc apart or NULL)
  const float* ln_gamma; const float* ln_beta; float ln_eps; const float* c_res;
  const float* c_bias;                                                             // tensor-core path only: [K] added per output column
  // tensor-core path only (ft.cu): pitch of the pre-split A rows (0: kpad(N)); B operand already split ([K][2 * kpad(N)] hi | lo
  // rows, as split_rows_kernel writes it); LayerNorm epilogue also / only emitting bf16 hi | lo rows (c may then be NULL);
  // window-ordered split output (see GemmTcParams::split_dst)
  int a_ld_pre; const void* b_pre;
  void* ln_hi; void* ln_lo; int ln_ld;
  void* split_dst[4]; const void* rowmap; int npix, n_img, img_shift;
};
int gemm_nt(const GemmNT& a, cudaStream_t st);

// out[i] = sum_b in[b*stride + i], i < n
int reduce_batch(const float* in, long long stride, float* out, int B, long long n, int accumulate, cudaStream_t st);
