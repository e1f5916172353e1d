// Attention backward on the tensor cores (tcgen05 / TMEM / TMA); see attn_bwd_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "attn_tc.cuh"

enum { ATTN_BWD_ROW = 0, ATTN_BWD_COL = 1, ATTN_BWD_PV = 2 };

// One launch = one gradient of  O = softmax(Q K^T / sqrt(128)) V  (rows r, columns c of a recomputed score tile):
//   ROW  (dQ): x = Q, y = K, g = dO, z = V   W[r,c] = e^{S - L[r]} (dP - D[r]),  out = W K / sqrt(128)
//   COL  (dK): x = K, y = Q, g = V,  z = dO  W[r,c] = e^{S - L[c]} (dP - D[c]),  out = W Q / sqrt(128)
//   PV   (dV): x = K, y = Q,         z = dO  W[r,c] = e^{S - L[c]},              out = W dO
// with S = x y^T / sqrt(128), dP = g z^T, L the row log-sum-exp of the forward and D = rowsum(dO o O).
struct AttnBwdTcArgs {
  const void* x_split;       // bf16 [nb][nr][256] token-major hi|lo
  const void* y_split;       // bf16 [nb][nc][256]
  const void* g_split;       // bf16 [nb][nr][256] (ROW, COL)
  const void* z_split;       // bf16 [nb][nc][256]
  const float* lse;          // [nb][n]: n = nr (ROW) or nc (COL, PV); with win.enabled both statistics are [B][h * w]
  const float* dsum;         // [nb][n]  (ROW, COL)
  float* out;                // ksplit == 1: [nb] x (nr * 128) in out_layout, batch stride out_stride_b
  long long out_stride_b;
  float* part;               // ksplit > 1: [ksplit][nb][nr * 128] partial sums (out_layout); add them with attn_bwd_tc_sum
  int mode, nb, nr, nc, out_layout;
  float sqrt_c;
  int ksplit;                // > 1: the column tiles of a row tile are dealt to ksplit CTAs
  AttnWinMap win;            // enabled: out is the [B][h][w][128] image, rows are scattered (needs ksplit == 1, NC layout)
};

int attn_bwd_tc(const AttnBwdTcArgs& a, cudaStream_t st);
int attn_bwd_tc_sum(const AttnBwdTcArgs& a, cudaStream_t st);
// D[b][r] = sum_c dO[b][r][c] O[b][r][c]; layout NC: [n][128] rows, CN: [128][n]; explicit batch strides (floats)
int attn_dsum(const float* d_o, long long do_stride_b, const float* o, long long o_stride_b, float* dsum, int nb, int n, int layout,
              cudaStream_t st);
