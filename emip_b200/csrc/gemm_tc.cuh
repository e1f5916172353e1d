// Tensor-core (tcgen05, 3-term split-bf16, fp32 accumulate) version of gemm_nn (gemm_simt.cuh); see conv_corr.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "gemm_simt.cuh"

// y[b][m][n] = sum_k W(b)[m][k] * X'(b)[k][n] (+ res), X' = optional LayerNorm of x over k -- the arguments of gemm_nn.
// Supported: no accumulate, K <= 1024, 16-byte aligned output rows.
bool gemm_nn_tc_supported(const GemmNN& a);
// scratch for the bf16 hi|lo operands: weights [nbw][M][Kp] x 2 and token-major activations [B][N][2*Kp]
size_t gemm_nn_tc_scratch_bytes(int B, int M, int K, int N, bool per_sample_w);
int gemm_nn_tc(const GemmNN& a, void* scratch, size_t scratch_bytes, cudaStream_t st);

// c[b][m][k] = sum_n A(b)[m][n] * B'(b)[k][n] -- the arguments of gemm_nt (one result per batch entry, no N split).
bool gemm_nt_tc_supported(const GemmNT& a);
size_t gemm_nt_tc_scratch_bytes(int B, int M, int K, int N);
int gemm_nt_tc(const GemmNT& a, void* scratch, size_t scratch_bytes, cudaStream_t st);
