// Tensor-core (tcgen05 / TMEM / TMA) backward of the pairwise-softmax family; see pair_bwd_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

// dX for problem p: rows = batch (x_base + p) of the split buffers, columns = batch (y_base + p).
// Per-row / per-column terms are indexed by p exactly as in PairBwdArgs (pair_common.cuh).
struct PairBwdTcArgs {
  const void* tok_split;     // bf16 [n_split][nr][256] token-major hi|lo: the row operand X
  const void* tok_split_y;   // bf16 [n_split][nc][256] token-major hi|lo: the column operand Y (S recompute)
  const void* chn_split_y;   // bf16 [n_split][256][ld] channel-major Y: rows 0..127 hi, 128..255 lo (dX = W Y)
  int n_split, x_base, y_base;
  const float* l1; const float* u; const float* u0; const float* t; long long t_stride_b;      // row-softmax term
  const float* l2; const float* w; const float* w0; const float* t2; long long t2_stride_b;    // column-softmax term
  const float* e; long long e_stride_b, e_stride_r, e_stride_c;                                 // direct term
  float* dx;                 // [nb][nr*128] in dx_layout
  int nb, nr, nc, dx_layout;
  float sqrt_c;
  int ksplit;                // > 1: split the key tiles over ksplit CTAs per row tile; dx then is [ksplit][nb][nr*128] partials
};

bool pair_bwd_tc_supported(int nr, int nc, int c);
long long pair_bwd_tc_chn_ld(int n);
size_t pair_bwd_tc_chn_bytes(int nb, int n);
// fp32 -> channel-major bf16 hi|lo (src2 != NULL: nb more batches appended); layout = layout of the source
int pair_bwd_tc_split_chn(const float* src, const float* src2, void* dst, int nb, int n, int layout, cudaStream_t st);
int pair_bwd_tc(const PairBwdTcArgs& a, cudaStream_t st);
