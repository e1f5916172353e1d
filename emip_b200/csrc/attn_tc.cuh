// Fused single-pass attention forward on the tensor cores (tcgen05 / TMEM / TMA); see attn_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

// Optional output scatter for window attention (window_attn.cu): problem p = block (p / B) of image (p % B), token t of
// a block of width bw is pixel (r0 + t / bw, c0 + t % bw) of the [B][h][w][128] output.
struct AttnWinMap {
  int enabled, B, h, w;
  int r0[16], c0[16], bw[16];
#ifdef __CUDACC__
  // index of token t of problem `prob` in the [B][h * w] pixel order (per-row statistics live in that order too)
  __device__ __forceinline__ size_t pixel(int prob, int t) const {
    const int blk = prob / B, img = prob - blk * B;
    const int wd = bw[blk];
    const int ty = t / wd, tx = t - ty * wd;
    return ((size_t)img * h + r0[blk] + ty) * w + c0[blk] + tx;
  }
#endif
};

// out[b][r][:] = softmax_c(Q[b][r] . K[b][c] / sqrt(128)) V[b][c][:]   for nb independent problems.
struct AttnTcArgs {
  const void* q_split;       // bf16 [nb][nq][256] token-major hi|lo (match_tc_split)
  const void* k_split;       // bf16 [nb][nk][256] token-major hi|lo
  const void* v_split;       // v_chn == 0: bf16 [nb][nk][256] token-major hi|lo values (read as an MN-major UMMA operand)
                             // v_chn == 1: bf16 [nb][256][ld] channel-major: rows 0..127 hi, 128..255 lo (pair_bwd_tc_split_chn)
  int v_chn;
  float* out;                // ksplit == 1: [nb] x (nq * 128) in out_layout, batch stride out_stride_b (floats)
  long long out_stride_b;
  float* lse;                // optional [nb][nq]: log-sum-exp of the scaled scores (natural log); [B][h * w] with win.enabled
  float* part_o;             // ksplit > 1: [ksplit][nb][nq * 128] un-normalised partial outputs (out_layout)
  float2* part_ml;           // ksplit > 1: [ksplit][nb][nq] (reference exponent in log2 units, partial row sum)
  int nb, nq, nk, out_layout;
  float sqrt_c;
  int ksplit;                // > 1: the key tiles of a row tile are dealt to ksplit CTAs (few row tiles, many keys)
  AttnWinMap win;            // enabled: out is the [B][h][w][128] image, rows are scattered (needs ksplit == 1, NC layout)
  // optional (ksplit == 1, NC layout): the output rows as bf16 hi / lo (rows out_ld elements apart, same row order as `out`) --
  // the pre-split A operand of the merge GEMM (ft.cu); `out` may then be NULL
  void* out_hi; void* out_lo; int out_ld;
  // optional: up to two MORE problem sets run by the same launch (window attention: the block groups of a shifted layer have
  // different token counts; as three launches each ended in its own partial wave -- at 8 pairs 64 + 128 + 64 items on 148 SMs).
  // Set i has its own operands and sizes; its problem p is block more[i].blk0 + p / win.B of image p % win.B.  Everything else is
  // shared.  Needs win.enabled, ksplit == 1, v_chn == 0; list the sets with the most key tiles first (items are dealt round-robin).
  int n_more;
  struct More { const void* q_split; const void* k_split; const void* v_split; int nb, nq, nk, blk0; } more[2];
};

bool attn_tc_supported(int nq, int nk, int c);
// diagnostics: the buffer set through emip_attn_tc_set_profile_buffer (NULL = off); shared with attn_bwd_tc.cu
unsigned long long* attn_tc_profile_buffer();
int attn_tc_fwd(const AttnTcArgs& a, cudaStream_t st);
// out / lse from the ksplit partials of attn_tc_fwd
int attn_tc_merge(const AttnTcArgs& a, cudaStream_t st);
