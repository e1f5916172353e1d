// Shared-memory tile helpers of the exact-fp32 CUDA-core kernels (pair_simt.cu, memory_read.cu):
// 64 x 64 score tiles over a 128-channel contraction, 256 threads as 16 x 16,
// thread (tx, ty) owns rows ty + 16a and columns tx + 16b (a, b < 4).
#pragma once
#include "common.cuh"

namespace simt {

constexpr int KC = 128;          // feature channels (contraction length)
constexpr int BM = 64;           // rows per CTA
constexpr int BN = 64;           // columns per tile
constexpr int LDX = KC + 4;      // padded smem row stride (floats) -> conflict-free LDS.128
constexpr int LDW = BN + 4;
constexpr int NT = 256;          // threads: 16 x 16, thread (tx,ty) owns rows ty+16a, cols tx+16b

// Load a [64 tokens x 128 channels] tile into smem as T[token][channel].
// layout 0: src is token-major [N][C]; layout 1: src is channel-major [C][N].
// Tokens >= n_valid are zero-filled.
__device__ __forceinline__ void load_tile(float* __restrict__ dst, const float* __restrict__ src, int layout,
                                          int tok0, int n_valid, int n_total) {
  const int tid = threadIdx.x;
  if (layout == 0) {
    // 64 rows x 32 float4; a warp reads one full 512 B row
    for (int i = tid; i < BM * (KC / 4); i += NT) {
      int r = i / (KC / 4), c4 = i % (KC / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tok0 + r < n_valid) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(tok0 + r) * KC) + c4);
      *reinterpret_cast<float4*>(dst + r * LDX + c4 * 4) = v;
    }
  } else {
    // channel-major: consecutive lanes read consecutive tokens of one channel
    for (int i = tid; i < BM * KC; i += NT) {
      int c = i / BM, r = i % BM;
      float v = 0.f;
      if (tok0 + r < n_valid) v = __ldg(src + (size_t)c * n_total + tok0 + r);
      dst[r * LDX + c] = v;
    }
  }
}

// acc[a][b] = sum_c Xs[ty+16a][c] * Ys[tx+16b][c]
__device__ __forceinline__ void s_tile(const float* __restrict__ Xs, const float* __restrict__ Ys, int tx, int ty,
                                       float (&acc)[4][4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
  for (int c4 = 0; c4 < KC / 4; ++c4) {
    float4 xa[4], yb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xa[a] = *reinterpret_cast<const float4*>(Xs + (ty + 16 * a) * LDX + c4 * 4);
#pragma unroll
    for (int b = 0; b < 4; ++b) yb[b] = *reinterpret_cast<const float4*>(Ys + (tx + 16 * b) * LDX + c4 * 4);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        acc[a][b] = fmaf(xa[a].x, yb[b].x, acc[a][b]);
        acc[a][b] = fmaf(xa[a].y, yb[b].y, acc[a][b]);
        acc[a][b] = fmaf(xa[a].z, yb[b].z, acc[a][b]);
        acc[a][b] = fmaf(xa[a].w, yb[b].w, acc[a][b]);
      }
  }
}

__device__ __forceinline__ float group16_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}


// dacc[a][0..7] += sum_c Ws[ty+16a][c] * Ys[c][tx*4..+3 | (tx+16)*4..+3]   (rows of Ys are tokens, 128 channels)
__device__ __forceinline__ void wy_accumulate(const float* __restrict__ Ws, const float* __restrict__ Ys, int tx, int ty,
                                              float (&dacc)[4][8]) {
#pragma unroll 2
  for (int c4 = 0; c4 < BN / 4; ++c4) {
    float4 wa[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) wa[a] = *reinterpret_cast<const float4*>(Ws + (ty + 16 * a) * LDW + c4 * 4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 y0 = *reinterpret_cast<const float4*>(Ys + (c4 * 4 + k) * LDX + tx * 4);
      float4 y1 = *reinterpret_cast<const float4*>(Ys + (c4 * 4 + k) * LDX + (tx + 16) * 4);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float wv = k == 0 ? wa[a].x : k == 1 ? wa[a].y : k == 2 ? wa[a].z : wa[a].w;
        dacc[a][0] = fmaf(wv, y0.x, dacc[a][0]);
        dacc[a][1] = fmaf(wv, y0.y, dacc[a][1]);
        dacc[a][2] = fmaf(wv, y0.z, dacc[a][2]);
        dacc[a][3] = fmaf(wv, y0.w, dacc[a][3]);
        dacc[a][4] = fmaf(wv, y1.x, dacc[a][4]);
        dacc[a][5] = fmaf(wv, y1.y, dacc[a][5]);
        dacc[a][6] = fmaf(wv, y1.z, dacc[a][6]);
        dacc[a][7] = fmaf(wv, y1.w, dacc[a][7]);
      }
    }
  }
}

// Writes a [64 rows x 128 channels] accumulator tile, scaled, to a channel-major global tensor
// dst[c * ld + row0 + r] (rows >= n_valid are dropped), staging through `stage` (>= BM * LDX floats).
// Callers must __syncthreads() before (stage may alias a tile that is still being read).
__device__ __forceinline__ void store_tile_cn(float* __restrict__ stage, const float (&dacc)[4][8], float scale,
                                              float* __restrict__ dst, long long ld, int row0, int n_valid, int tx, int ty,
                                              bool accumulate) {
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float* d = stage + (ty + 16 * a) * LDX;
    *reinterpret_cast<float4*>(d + tx * 4) =
        make_float4(dacc[a][0] * scale, dacc[a][1] * scale, dacc[a][2] * scale, dacc[a][3] * scale);
    *reinterpret_cast<float4*>(d + (tx + 16) * 4) =
        make_float4(dacc[a][4] * scale, dacc[a][5] * scale, dacc[a][6] * scale, dacc[a][7] * scale);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BM * KC; i += NT) {
    int c = i / BM, r = i % BM;
    int row = row0 + r;
    if (row < n_valid) {
      float* p = dst + (size_t)c * ld + row;
      float v = stage[r * LDX + c];
      *p = accumulate ? (*p + v) : v;
    }
  }
}

}  // namespace simt
