// Tensor-core GEMMs on bf16 hi|lo split operands (tcgen05 SS-mode UMMA, fp32 accumulation in TMEM, TMA-fed 128-byte-swizzled
// K-major tiles): D = A.hi B.hi + A.lo B.hi + A.hi B.lo.  One kernel, three operand / epilogue modes (gemm_tc.cu):
//   mode 0  G = A B^T * scale, written as bf16 hi | lo                       (conv_corr: weight x features)
//   mode 1  out = conv3x3 as a GEMM whose B tiles are shifted 4-D TMA boxes   (conv_corr: taps of f0, zero fill = padding)
//   mode 2  y = A Bt^T (+ residual), fp32 row-major                            (gemm_nn_tc / gemm_nt_tc: the Injector)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stddef.h>
#include "gemm_simt.cuh"

struct GemmTcParams {
  int mode;                 // see above
  int M;                    // rows of A / of the output
  int n_mtiles, n_ntiles;   // tiles per batch entry
  int total_tiles;          // batch * n_mtiles * n_ntiles (set by gemm_tc_launch)
  int n_tile;               // UMMA N (multiple of 16, <= 256)
  int kchunks;              // K / 64 (rounded up; the TMA unit zero-fills the tail)
  int stages, stage_bytes, b_bytes;
  int W, H, R;              // mode 1: image geometry, image rows per pixel tile
  int cpt, lo_off;          // mode 1: K chunks per tap (input channels / 64, rounded up) and offset of the lo half in a token row
  int a_shared;             // mode 1: A is one weight matrix for every batch entry (0: one per sample, conv_corr's G[b])
  float scale;              // mode 0
  const float* bias;        // mode 1 (and per-row bias of mode 2)
  const float* ep_scale;    // mode 1: out = act(acc * ep_scale[row] + bias[row]) (NULL: scale 1)
  int ep_relu;              // mode 1: act = ReLU
  // mode 1, tok_hi != NULL: the tile goes out as TOKEN-MAJOR bf16 hi | lo rows [b][pixel][hi tok_lo_off | lo] of pitch tok_ld --
  // the input operand of a following emip_conv3x3 (conv_corr[0] + BN + ReLU -> conv_corr[3]) -- instead of fp32 [b][o][pixel];
  // rows >= M (the channel padding up to a multiple of 128) are written as zeros
  __nv_bfloat16* tok_hi; int tok_ld, tok_lo_off;
  __nv_bfloat16* g_hi;      // mode 0 output: [B][M][128] hi, lo
  __nv_bfloat16* g_lo;
  float* out;               // mode 1 output: [B][M][H*W]
  // mode 2: y[b][m][n] = sum_k A(b)[m][k] Bt[b][n][k] (+ res): A = weights (shared: a_batched = 0), Bt = token-major
  // activations [B][N][2*Kp] (hi | lo); y / res rows are ldy / ldr floats apart
  int a_batched, Kp, N;
  int ksplit;               // mode 2, K-major B: > 1 = the batch index is (sample, K split), kchunks = chunks per split
  int b_mn, Np;             // b_mn: Bt is [B][K][2*Np] (k rows, n contiguous, hi | lo) and is read as an MN-major operand
  float* y; long long y_stride_b; int ldy;
  const float* res; long long res_stride_b; int ldr;
  // mode 2, out_split != 0: the tile is written as bf16 hi | lo rows of pitch ldg into g_hi / g_lo (the A operand of a
  // following GEMM) instead of fp32 y; out_split == 2 applies the exact GELU first.  Batch 1, N % 32 == 0.
  int out_split, ldg;
  // mode 2, ln_gamma != NULL (N == n_tile == 128, one output row per thread): y = res + LayerNorm_N(tile) * gamma + beta
  const float* ln_gamma; const float* ln_beta; float ln_eps;
  const float* bias_n;      // mode 2, plain fp32 epilogue: per-column bias [N] (NULL: none)
  // mode 2, LayerNorm epilogue: the normalised row also as bf16 hi | lo (rows ln_ld elements apart) -- the pre-split A operand
  // of the next GEMM; y may then be NULL (only the split is wanted)
  __nv_bfloat16* ln_hi; __nv_bfloat16* ln_lo; int ln_ld;
  // mode 2, out_split == 3 (n_tile == 128, batch 1): column tile nt goes to the bf16 buffer split_dst[nt] whose rows are
  // [hi 128 | lo 128]; row r = (image, pixel) lands in row rowmap[pixel].x + ((image + img_shift) % n_img) * rowmap[pixel].y --
  // the window-ordered q / k / v operands of the FeatureTransformer attention (ft.cu), no separate gather / split pass
  __nv_bfloat16* split_dst[4]; const int2* rowmap; int npix, n_img, img_shift;
  unsigned long long* prof; // diagnostics (tools/gemm_roles.py): [grid][8] cycles: producer total, wait empty | issuer total, wait full, wait acc_empty | epilogue warp 0 total, wait acc_full, tiles
  int dbg;                  // diagnostics (emip_debug_gemm_wide_tiles): bit 0 = epilogue warps skip their work, bit 1 = producer skips the TMA loads
  int epi_stage;            // mode 2: per-warp shared-memory staging of the epilogue (set by gemm_tc_launch): coalesced row stores
};

constexpr int GEMM_TC_KCH = 64;                    // K elements per chunk (one 128-byte swizzle row of bf16)
constexpr int GEMM_TC_A_BYTES = 128 * 128;         // one [128 x 64] bf16 operand tile

// diagnostics (emip_debug_gemm_wide_tiles bit 2): the FeatureTransformer's FFN as two gemm_tc launches (A / B runs, parity tests)
int gemm_tc_debug_two_launch_mlp();
// bf16 tiled tensor map with 128-byte swizzle (rank 3 or 4), out-of-bounds elements read as zero
int gemm_tc_make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box);
// launches grid = batch * n_mtiles * n_ntiles CTAs; fills stages / stage_bytes from n_tile when stages == 0
int gemm_tc_launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b, GemmTcParams p, int batch,
                   cudaStream_t st);

// y[b][m][n] = sum_k W(b)[m][k] * X'(b)[k][n] (+ res), X' = optional LayerNorm of x over k -- the arguments of gemm_nn.
// Supported: no accumulate, K <= 1024, 16-byte aligned output rows.
bool gemm_nn_tc_supported(const GemmNN& a);
// scratch for the bf16 hi|lo operands: weights [nbw][M][Kp] x 2 and token-major activations [B][N][2*Kp]
size_t gemm_nn_tc_scratch_bytes(int B, int M, int K, int N, bool per_sample_w);
int gemm_nn_tc(const GemmNN& a, void* scratch, size_t scratch_bytes, cudaStream_t st);
// where gemm_nn_tc expects the activation operand inside its scratch: bf16 [B][K][2 * Np] (hi | lo per row, zero padded to
// Np = *np_out); a producer may write it there itself and set GemmNN::x_presplit
void* gemm_nn_tc_act_operand(void* scratch, int B, int M, int K, int N, bool per_sample_w, int* np_out);

// c[b][m][k] = sum_n A(b)[m][n] * B'(b)[k][n] -- the arguments of gemm_nt (one result per batch entry, no N split).
bool gemm_nt_tc_supported(const GemmNT& a);
size_t gemm_nt_tc_scratch_bytes(int B, int M, int K, int N);
int gemm_tc_split_rows(const float* x, int M, int K, void* hi, void* lo, cudaStream_t st);   // fp32 rows -> bf16 hi, lo [M][kpad(K)]
size_t gemm_nt_tc_scratch_bytes_presplit(int K, int N);
// weight [K][N] fp32 -> GemmNT::b_pre format (bf16 [K][2 * kpad(N)], hi | lo per row), gemm_tc_split_b_bytes(K, N) bytes
int gemm_tc_split_b(const float* w, int ldb, int K, int N, void* dst, cudaStream_t st);
size_t gemm_tc_split_b_bytes(int K, int N);    // GemmNT::a_hi_pre set: only the B operand is split
// nsplit > 1: the contraction axis is dealt to nsplit CTAs per tile; c then holds B * nsplit partial matrices
// (c_stride_b apart, partial (b, s) at index b * nsplit + s) for the caller's batch reduction
int gemm_nt_tc(const GemmNT& a, void* scratch, size_t scratch_bytes, cudaStream_t st, int nsplit = 1);

// f1 backward (gemm_tc.cu): df0, df1 [B][128][H*W], dweight [O][H*W][3][3], dbias [O] (may be NULL) from dout [B][O][H*W];
// w_prep = the forward's prepared weight (hi | lo rows of pitch w_prep_ld); scratch 1024-byte aligned
size_t conv_corr_bwd_scratch_bytes(int B, int O, int P);
int conv_corr_bwd_tc(const float* f0, const float* f1, const float* weight, const void* w_prep, long long w_prep_ld,
                     const float* dout, float* df0, float* df1, float* dweight, float* dbias, void* scratch, size_t scratch_bytes,
                     int B, int H, int W, int O, cudaStream_t st);
