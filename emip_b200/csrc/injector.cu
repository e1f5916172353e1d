// K4: prompt fusion (camouflaged feeder / motion collector) -- forward and backward.
//
// Replaces reference model/EMIP_short/motion/PromptInteract.py:452-464 (Injector) =
// TransformerBlock_MDTA :436-450 = x + MDTA(LN1(x), LN2(x1)), then x + GDFN(LN3(x)) with
// Attention_MDTA :390-432, FeedForward :367-385, WithBias_LayerNorm :333-349, and the autograd
// graph torch derives for them.  dim 128, 2 heads x 64 channels, hidden 340, all convs bias-free.
//
// Everything is exact fp32 on the CUDA cores: the injector's inputs/outputs have to stay
// fp32-accurate (bf16 I/O here costs 4.7e-2 end to end), and at ~1 MB per tensor
// the block is launch/HBM bound, not FLOP bound.  What the reference runs as ~25 library
// launches with three `rearrange` copies becomes 10 launches (forward):
//   ln_stats x2 | gemm (LN fused on load) x2 | depthwise 3x3 (+ row sum-of-squares) x2 |
//   64x64 Gram per (sample, head) | softmax + fold attn into project_out | gemm + residual |
//   ln_stats | gemm (LN on load) | depthwise 3x3 + gelu gate | gemm + residual
// Layout: activations stay [B, C, N] (NCHW, N = H*W contiguous): no permutes anywhere.
// All reductions are deterministic (two-stage, no atomics).
#include "common.cuh"
#include "../../include/emip_b200.h"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include <math.h>

namespace {

constexpr int DIM = 128;
constexpr int HEADS = 2;
constexpr int HD = 64;          // channels per head
constexpr int HID = 340;        // int(128 * 2.66)
constexpr int HID2 = 680;
constexpr float LN_EPS = 1e-5f;
constexpr float NORM_EPS = 1e-12f;

// ------------------------------------------------------------------ LayerNorm statistics
// mean / rstd over the C channels of every pixel (PromptInteract.py:346-349: biased variance, eps 1e-5)
// Block = 32 pixels x 8 channel groups (C = 128: 16 channels per thread, all loads in flight, values kept in registers
// for the second pass); the groups are combined through shared memory in a fixed order.  r1: one thread per pixel
// walking 2 x 128 strided loads filled a tenth of the chip (15 us for a 16 MB read).
constexpr int LN_PX = 32, LN_GR = 8, LN_CPT = DIM / LN_GR;
__global__ void __launch_bounds__(LN_PX * LN_GR)
ln_stats_kernel(const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ rstd, int C, int N,
                const float* __restrict__ x2 = nullptr, float* __restrict__ mean2 = nullptr, float* __restrict__ rstd2 = nullptr) {
  __shared__ float red[LN_GR][LN_PX];
  if (blockIdx.z == 1) { x = x2; mean = mean2; rstd = rstd2; }      // gridDim.z == 2: the statistics of two tensors in one launch
  const int b = blockIdx.y, px = threadIdx.x & (LN_PX - 1), gr = threadIdx.x / LN_PX;
  const int n = blockIdx.x * LN_PX + px;
  const bool ok = n < N;
  const float* xp = x + (size_t)b * C * N + (size_t)gr * LN_CPT * N + n;
  float v[LN_CPT];
#pragma unroll
  for (int j = 0; j < LN_CPT; ++j) v[j] = ok ? __ldg(xp + (size_t)j * N) : 0.f;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_CPT; ++j) s += v[j];
  red[gr][px] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int g = 0; g < LN_GR; ++g) tot += red[g][px];
  const float mu = tot / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < LN_CPT; ++j) {
    const float d = v[j] - mu;
    q = fmaf(d, d, q);
  }
  __syncthreads();
  red[gr][px] = q;
  __syncthreads();
  if (gr == 0 && ok) {
    float var = 0.f;
#pragma unroll
    for (int g = 0; g < LN_GR; ++g) var += red[g][px];
    mean[(size_t)b * N + n] = mu;
    rstd[(size_t)b * N + n] = 1.0f / sqrtf(var / (float)C + LN_EPS);
  }
}

// ------------------------------------------------------------------ GEMM building blocks (gemm_simt.cuh)
constexpr int GB = 64;   // block tile (both dims)
constexpr int GK = 16;   // contraction step

__device__ __forceinline__ float ln_apply(float v, int k, size_t bn, const GemmNN& a) {
  return (v - __ldg(a.mean + bn)) * __ldg(a.rstd + bn) * __ldg(a.gamma + k) + __ldg(a.beta + k);
}

// Y[b][m][n] = sum_k W(b)[m][k] X'(b)[k][n] (+res) ; 256 threads, 64x64 tile, 4x4 outputs per thread
__global__ void __launch_bounds__(256) gemm_nn_kernel(GemmNN a) {
  __shared__ float Ws[GK][GB + 4];
  __shared__ float Xs[GK][GB + 4];
  const int b = blockIdx.z, m0 = blockIdx.y * GB, n0 = blockIdx.x * GB;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* W = a.w + (size_t)b * a.w_stride_b;
  const float* X = a.x + (size_t)b * a.x_stride_b;
  const bool ln = a.mean != nullptr;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < a.K; k0 += GK) {
    // W tile: element (m, k)
    for (int i = threadIdx.x; i < GB * GK; i += 256) {
      int m, k;
      if (a.w_trans) { k = i / GB; m = i % GB; } else { m = i / GK; k = i % GK; }
      float v = 0.f;
      if (m0 + m < a.M && k0 + k < a.K)
        v = __ldg(W + (a.w_trans ? (size_t)(k0 + k) * a.ldw + (m0 + m) : (size_t)(m0 + m) * a.ldw + (k0 + k)));
      Ws[k][m] = v;
    }
    // X tile: element (k, n), n contiguous
    for (int i = threadIdx.x; i < GK * GB; i += 256) {
      const int k = i / GB, n = i % GB;
      float v = 0.f;
      if (k0 + k < a.K && n0 + n < a.N) {
        v = __ldg(X + (size_t)(k0 + k) * a.ldx + (n0 + n));
        if (ln) v = ln_apply(v, k0 + k, (size_t)b * a.N + n0 + n, a);
      }
      Xs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[k][ty * 4]);
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[k][tx * 4]);
      const float wa[4] = {wv.x, wv.y, wv.z, wv.w}, xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wa[i], xa[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* Y = a.y + (size_t)b * a.y_stride_b;
  const float* R = a.res ? a.res + (size_t)b * a.res_stride_b : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j];
      if (R) v += __ldg(R + (size_t)m * a.ldr + n);
      float* yp = Y + (size_t)m * a.ldy + n;
      *yp = a.accumulate ? (*yp + v) : v;
    }
  }
}

// C[b*S+s][m][k] = sum_{n in split s} A(b)[m][n] B'(b)[k][n]
struct GemmNTk {
  GemmNT g;
  int nsplit, chunk;
};
__global__ void __launch_bounds__(256) gemm_nt_kernel(GemmNTk p) {
  const GemmNT& a = p.g;
  __shared__ float As[32][GB + 4];
  __shared__ float Bs[32][GB + 4];
  const int bs = blockIdx.z, b = bs / p.nsplit, s = bs % p.nsplit;
  const int m0 = blockIdx.y * GB, k0 = blockIdx.x * GB;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* A = a.a + (size_t)b * a.a_stride_b;
  const float* Bm = a.bm + (size_t)b * a.b_stride_b;
  const bool ln = a.mean != nullptr;
  const int nbeg = s * p.chunk, nend = min(a.N, nbeg + p.chunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int n0 = nbeg; n0 < nend; n0 += 32) {
    for (int i = threadIdx.x; i < GB * 32; i += 256) {
      const int r = i / 32, n = i % 32;          // 32 consecutive n per row: one 128-byte segment per warp
      float va = 0.f, vb = 0.f;
      if (n0 + n < nend) {
        if (m0 + r < a.M) va = __ldg(A + (size_t)(m0 + r) * a.lda + n0 + n);
        if (k0 + r < a.K) {
          vb = __ldg(Bm + (size_t)(k0 + r) * a.ldb + n0 + n);
          if (ln) {
            const size_t bn = (size_t)b * a.N + n0 + n;
            vb = (vb - __ldg(a.mean + bn)) * __ldg(a.rstd + bn) * __ldg(a.gamma + k0 + r) + __ldg(a.beta + k0 + r);
          }
        }
      }
      As[n][r] = va;
      Bs[n][r] = vb;
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < 32; ++n) {
      const float4 av = *reinterpret_cast<const float4*>(&As[n][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[n][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, ba[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], ba[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* C = a.c + (size_t)bs * a.c_stride_b;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < a.K) C[(size_t)m * a.ldc + k] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------ 128 x 128 tile SGEMM (fast paths)
// 256 threads as 16 x 16; thread (tx, ty) owns rows {4ty..4ty+3, 64+4ty..} and columns {4tx.., 64+4tx..} (8 x 8
// outputs), operands staged k-major in double-buffered shared memory with register prefetch of the next slab.
// The first version of these GEMMs (64 x 64 tiles, 4 x 4 per thread, no prefetch) reached 16 TFLOP/s and was 60 %
// of the injector's forward time.
constexpr int FB = 128;   // tile (both dims)
constexpr int FK = 8;     // contraction slab

__device__ __forceinline__ void fma_slab(const float (*As)[FB], const float (*Bs)[FB], int tx, int ty, float (&acc)[8][8]) {
#pragma unroll
  for (int k = 0; k < FK; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
    const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
    const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// Y[b][m][n] = sum_k W(b)[m][k] X'(b)[k][n] (+res): requires N % 4 == 0, ldx/ldy/ldr % 4 == 0, 16-byte aligned bases.
__global__ void __launch_bounds__(256, 2) gemm_nn_fast_kernel(GemmNN a) {
  __shared__ __align__(16) float As[2][FK][FB];
  __shared__ __align__(16) float Bs[2][FK][FB];
  const int b = blockIdx.z, m0 = blockIdx.y * FB, n0 = blockIdx.x * FB;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* W = a.w + (size_t)b * a.w_stride_b;
  const float* X = a.x + (size_t)b * a.x_stride_b;
  const bool ln = a.mean != nullptr;
  // B slab: 8 rows x 32 float4; this thread always loads row kb, columns nb4..nb4+3
  const int kb = tid >> 5, nb4 = n0 + (tid & 31) * 4;
  const bool n_ok = nb4 < a.N;
  float4 mu4 = make_float4(0.f, 0.f, 0.f, 0.f), rs4 = make_float4(1.f, 1.f, 1.f, 1.f);
  if (ln && n_ok) {
    mu4 = __ldg(reinterpret_cast<const float4*>(a.mean + (size_t)b * a.N + nb4));
    rs4 = __ldg(reinterpret_cast<const float4*>(a.rstd + (size_t)b * a.N + nb4));
  }
  // A slab: 128 rows x 8 k.  w_trans: 8 k-rows x 32 float4 along m; else thread -> (row tid/2, four k at (tid%2)*4)
  const int am = a.w_trans ? (tid & 31) * 4 : (tid >> 1), ak = a.w_trans ? (tid >> 5) : (tid & 1) * 4;
  float4 ra, rb;
  auto load = [&](int k0) {
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.w_trans) {
      if (k0 + ak < a.K && m0 + am < a.M) ra = __ldg(reinterpret_cast<const float4*>(W + (size_t)(k0 + ak) * a.ldw + m0 + am));
    } else {
      if (m0 + am < a.M) {
        const float* wp = W + (size_t)(m0 + am) * a.ldw + k0 + ak;
        if (k0 + ak + 3 < a.K) ra = __ldg(reinterpret_cast<const float4*>(wp));
        else {
          if (k0 + ak < a.K) ra.x = __ldg(wp);
          if (k0 + ak + 1 < a.K) ra.y = __ldg(wp + 1);
          if (k0 + ak + 2 < a.K) ra.z = __ldg(wp + 2);
        }
      }
    }
    if (k0 + kb < a.K && n_ok) {
      rb = __ldg(reinterpret_cast<const float4*>(X + (size_t)(k0 + kb) * a.ldx + nb4));
      if (ln) {
        const float gk = __ldg(a.gamma + k0 + kb), bk = __ldg(a.beta + k0 + kb);
        rb.x = (rb.x - mu4.x) * rs4.x * gk + bk;
        rb.y = (rb.y - mu4.y) * rs4.y * gk + bk;
        rb.z = (rb.z - mu4.z) * rs4.z * gk + bk;
        rb.w = (rb.w - mu4.w) * rs4.w * gk + bk;
      }
    }
  };
  auto store = [&](int s) {
    if (a.w_trans) *reinterpret_cast<float4*>(&As[s][ak][am]) = ra;
    else { As[s][ak][am] = ra.x; As[s][ak + 1][am] = ra.y; As[s][ak + 2][am] = ra.z; As[s][ak + 3][am] = ra.w; }
    *reinterpret_cast<float4*>(&Bs[s][kb][(tid & 31) * 4]) = rb;
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  load(0);
  store(0);
  __syncthreads();
  int s = 0;
  for (int k0 = 0; k0 < a.K; k0 += FK) {
    const bool more = k0 + FK < a.K;
    if (more) load(k0 + FK);                      // global loads of the next slab fly during the FMAs
    fma_slab(As[s], Bs[s], tx, ty, acc);
    if (more) store(s ^ 1);
    __syncthreads();
    s ^= 1;
  }
  float* Y = a.y + (size_t)b * a.y_stride_b;
  const float* R = a.res ? a.res + (size_t)b * a.res_stride_b : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= a.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      if (n >= a.N) continue;
      float4 v = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
      if (R) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(R + (size_t)m * a.ldr + n));
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      float4* yp = reinterpret_cast<float4*>(Y + (size_t)m * a.ldy + n);
      if (a.accumulate) { const float4 o = *yp; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      *yp = v;
    }
  }
}

// C[b*S+s][m][k] = sum_{n in split s} A(b)[m][n] B'(b)[k][n]: both operands contiguous along the contraction.
// Requires lda/ldb % 4 == 0, N % 4 == 0, chunk % 4 == 0, 16-byte aligned bases.
constexpr int TK = 16;    // contraction slab of the NT kernel (one float4 per thread and operand)
__global__ void __launch_bounds__(256, 2) gemm_nt_fast_kernel(GemmNTk p) {
  const GemmNT& a = p.g;
  __shared__ __align__(16) float As[2][TK][FB + 4];
  __shared__ __align__(16) float Bs[2][TK][FB + 4];
  const int bs = blockIdx.z, b = bs / p.nsplit, sp = bs % p.nsplit;
  const int m0 = blockIdx.y * FB, k0 = blockIdx.x * FB;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* A = a.a + (size_t)b * a.a_stride_b;
  const float* Bm = a.bm + (size_t)b * a.b_stride_b;
  const bool ln = a.mean != nullptr;
  const int nbeg = sp * p.chunk, nend = min(a.N, nbeg + p.chunk);
  // slab: 128 rows x 16 n = 512 float4 per operand: thread -> rows r0 and r0 + 64, four n at n4
  const int r0 = tid >> 2, n4 = (tid & 3) * 4;
  float gam[2] = {0.f, 0.f}, bet[2] = {0.f, 0.f};
  if (ln) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (k0 + r0 + 64 * h < a.K) { gam[h] = __ldg(a.gamma + k0 + r0 + 64 * h); bet[h] = __ldg(a.beta + k0 + r0 + 64 * h); }
  }
  float4 ra[2], rb[2];
  auto load = [&](int n0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int n = n0 + n4;
      if (n < nend) {
        if (m0 + r0 + 64 * h < a.M) ra[h] = __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + r0 + 64 * h) * a.lda + n));
        if (k0 + r0 + 64 * h < a.K) {
          rb[h] = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)(k0 + r0 + 64 * h) * a.ldb + n));
          if (ln) {
            const float4 mu = __ldg(reinterpret_cast<const float4*>(a.mean + (size_t)b * a.N + n));
            const float4 rs = __ldg(reinterpret_cast<const float4*>(a.rstd + (size_t)b * a.N + n));
            rb[h].x = (rb[h].x - mu.x) * rs.x * gam[h] + bet[h];
            rb[h].y = (rb[h].y - mu.y) * rs.y * gam[h] + bet[h];
            rb[h].z = (rb[h].z - mu.z) * rs.z * gam[h] + bet[h];
            rb[h].w = (rb[h].w - mu.w) * rs.w * gam[h] + bet[h];
          }
        }
      }
    }
  };
  auto store = [&](int s) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = r0 + 64 * h;
      As[s][n4][r] = ra[h].x; As[s][n4 + 1][r] = ra[h].y; As[s][n4 + 2][r] = ra[h].z; As[s][n4 + 3][r] = ra[h].w;
      Bs[s][n4][r] = rb[h].x; Bs[s][n4 + 1][r] = rb[h].y; Bs[s][n4 + 2][r] = rb[h].z; Bs[s][n4 + 3][r] = rb[h].w;
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (nbeg < nend) {
    load(nbeg);
    store(0);
    __syncthreads();
    int s = 0;
    for (int n0 = nbeg; n0 < nend; n0 += TK) {
      const bool more = n0 + TK < nend;
      if (more) load(n0 + TK);
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[s][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[s][k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[s][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[s][k][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (more) store(s ^ 1);
      __syncthreads();
      s ^= 1;
    }
  }
  float* C = a.c + (size_t)bs * a.c_stride_b;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (k < a.K) C[(size_t)m * a.ldc + k] = acc[i][j];
    }
  }
}

// out[b][i] = sum_s part[(b*S + s)][i]
__global__ void sum_splits_kernel(const float* __restrict__ part, float* __restrict__ out, int nb, int S, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nb * n) return;
  const int b = idx / n, i = idx % n;
  float acc = 0.f;
  for (int s = 0; s < S; ++s) acc += __ldg(part + ((size_t)b * S + s) * n + i);
  out[idx] = acc;
}

__global__ void reduce_batch_kernel(const float* __restrict__ in, long long stride, float* __restrict__ out, int B,
                                    long long n, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += __ldg(in + (size_t)b * stride + i);   // fixed order: deterministic
  out[i] = accumulate ? out[i] + s : s;
}

// ------------------------------------------------------------------ depthwise 3x3, one (sample, channel) plane per CTA
// The whole plane lives in shared memory with a zero halo, so forward, gelu gate, backward-data and
// backward-weight are all CTA-local.
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}

__device__ __forceinline__ void load_plane(float* s, const float* __restrict__ g, int H, int W) {
  const int Wp = W + 2, tot = (H + 2) * Wp;
  for (int i = threadIdx.x; i < tot; i += blockDim.x) {
    const int y = i / Wp - 1, x = i % Wp - 1;
    s[i] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(g + y * W + x) : 0.f;
  }
}
// cross-correlation with the 3x3 kernel k (nn.Conv2d semantics, padding 1)
__device__ __forceinline__ float conv9(const float* s, int Wp, int y, int x, const float* k) {
  const float* p = s + y * Wp + x;            // top-left of the window in the padded plane
  return p[0] * k[0] + p[1] * k[1] + p[2] * k[2] + p[Wp] * k[3] + p[Wp + 1] * k[4] + p[Wp + 2] * k[5] +
         p[2 * Wp] * k[6] + p[2 * Wp + 1] * k[7] + p[2 * Wp + 2] * k[8];
}
// transposed (backward-data) convolution: d_in(y,x) = sum_{ky,kx} d_out(y-ky+1, x-kx+1) k[ky][kx]
__device__ __forceinline__ float conv9_flipped(const float* s, int Wp, int y, int x, const float* k) {
  const float* p = s + y * Wp + x;
  return p[0] * k[8] + p[1] * k[7] + p[2] * k[6] + p[Wp] * k[5] + p[Wp + 1] * k[4] + p[Wp + 2] * k[3] +
         p[2 * Wp] * k[2] + p[2 * Wp + 1] * k[1] + p[2 * Wp + 2] * k[0];
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];   // fixed order
  return s;
}

// out(c) = dw3x3(in(c)); channels >= split go to out2 (channel c - split): splits kv into k and v.
// sumsq (optional): per (b, c) sum of squares of the output (for F.normalize over the pixels).
__global__ void __launch_bounds__(256)
dwconv_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out, float* __restrict__ out2,
                  float* __restrict__ sumsq, int C, int split, int H, int W, int sq_C) {
  extern __shared__ float sm[];
  __shared__ float red[8];
  const int c = blockIdx.x, b = blockIdx.y, N = H * W, Wp = W + 2;
  load_plane(sm, in + ((size_t)b * C + c) * N, H, W);
  float k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
  __syncthreads();
  float* o = (c < split) ? out + ((size_t)b * split + c) * N : out2 + ((size_t)b * (C - split) + (c - split)) * N;
  float ss = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float v = conv9(sm, Wp, i / W, i % W, k);
    o[i] = v;
    ss = fmaf(v, v, ss);
  }
  if (sumsq != nullptr && c < sq_C) {                     // block-uniform: |row|^2 of the first sq_C channels, [B][sq_C]
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) sumsq[(size_t)b * sq_C + c] = ss;
  }
}

// backward of the above: dout(c) comes from dout (c < split) or dout2; din = flipped conv; dw partial per (b, c)
__global__ void __launch_bounds__(256)
dwconv_bwd_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ dout,
                  const float* __restrict__ dout2, float* __restrict__ din, float* __restrict__ dw_part, int C, int split,
                  int H, int W) {
  extern __shared__ float sm[];
  __shared__ float red[8];
  const int c = blockIdx.x, b = blockIdx.y, N = H * W, Wp = W + 2, P = (H + 2) * Wp;
  float* s_in = sm;
  float* s_do = sm + P;
  load_plane(s_in, in + ((size_t)b * C + c) * N, H, W);
  const float* dg = (c < split) ? dout + ((size_t)b * split + c) * N : dout2 + ((size_t)b * (C - split) + (c - split)) * N;
  load_plane(s_do, dg, H, W);
  float k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
  __syncthreads();
  float dk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) dk[i] = 0.f;
  float* di = din + ((size_t)b * C + c) * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int y = i / W, x = i % W;
    di[i] = conv9_flipped(s_do, Wp, y, x, k);
    const float g = s_do[(y + 1) * Wp + x + 1];
    const float* p = s_in + y * Wp + x;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) dk[ky * 3 + kx] = fmaf(g, p[ky * Wp + kx], dk[ky * 3 + kx]);
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float s = block_sum(dk[i], red);
    if (threadIdx.x == 0) dw_part[((size_t)b * C + c) * 9 + i] = s;
  }
}

// GDFN gate, forward: g(c) = gelu(dw(t_pre[c])) * dw(t_pre[c + HID])      (PromptInteract.py:382-383)
__global__ void __launch_bounds__(256)
gdfn_gate_fwd_kernel(const float* __restrict__ tpre, const float* __restrict__ w, float* __restrict__ g, int H, int W) {
  extern __shared__ float sm[];
  const int c = blockIdx.x, b = blockIdx.y, N = H * W, Wp = W + 2, P = (H + 2) * Wp;
  load_plane(sm, tpre + ((size_t)b * HID2 + c) * N, H, W);
  load_plane(sm + P, tpre + ((size_t)b * HID2 + c + HID) * N, H, W);
  float k1[9], k2[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { k1[i] = __ldg(w + c * 9 + i); k2[i] = __ldg(w + (c + HID) * 9 + i); }
  __syncthreads();
  float* o = g + ((size_t)b * HID + c) * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int y = i / W, x = i % W;
    o[i] = gelu_erf(conv9(sm, Wp, y, x, k1)) * conv9(sm + P, Wp, y, x, k2);
  }
}

// GDFN gate, backward: recomputes t1, t2 from t_pre, forms dt1 = dg t2 gelu'(t1), dt2 = dg gelu(t1), then the
// depthwise backward (data + weight) for both channels.  Also re-emits g for the project_out weight gradient.
__global__ void __launch_bounds__(256)
gdfn_gate_bwd_kernel(const float* __restrict__ tpre, const float* __restrict__ w, const float* __restrict__ dg,
                     float* __restrict__ dtpre, float* __restrict__ g_out, float* __restrict__ dw_part, int H, int W) {
  extern __shared__ float sm[];
  __shared__ float red[8];
  const int c = blockIdx.x, b = blockIdx.y, N = H * W, Wp = W + 2, P = (H + 2) * Wp;
  float* s1 = sm;            // t_pre[c]
  float* s2 = sm + P;        // t_pre[c + HID]
  float* d1 = sm + 2 * P;    // dt1 (padded)
  float* d2 = sm + 3 * P;    // dt2
  load_plane(s1, tpre + ((size_t)b * HID2 + c) * N, H, W);
  load_plane(s2, tpre + ((size_t)b * HID2 + c + HID) * N, H, W);
  for (int i = threadIdx.x; i < P; i += blockDim.x) { d1[i] = 0.f; d2[i] = 0.f; }
  float k1[9], k2[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { k1[i] = __ldg(w + c * 9 + i); k2[i] = __ldg(w + (c + HID) * 9 + i); }
  __syncthreads();
  const float* dgp = dg + ((size_t)b * HID + c) * N;
  float* go = g_out + ((size_t)b * HID + c) * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int y = i / W, x = i % W;
    const float t1 = conv9(s1, Wp, y, x, k1), t2 = conv9(s2, Wp, y, x, k2);
    const float gg = gelu_erf(t1), d = __ldg(dgp + i);
    go[i] = gg * t2;
    d1[(y + 1) * Wp + x + 1] = d * t2 * gelu_erf_grad(t1);
    d2[(y + 1) * Wp + x + 1] = d * gg;
  }
  __syncthreads();
  float dk1[9], dk2[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { dk1[i] = 0.f; dk2[i] = 0.f; }
  float* o1 = dtpre + ((size_t)b * HID2 + c) * N;
  float* o2 = dtpre + ((size_t)b * HID2 + c + HID) * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int y = i / W, x = i % W;
    o1[i] = conv9_flipped(d1, Wp, y, x, k1);
    o2[i] = conv9_flipped(d2, Wp, y, x, k2);
    const float g1 = d1[(y + 1) * Wp + x + 1], g2 = d2[(y + 1) * Wp + x + 1];
    const float* p1 = s1 + y * Wp + x;
    const float* p2 = s2 + y * Wp + x;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        dk1[ky * 3 + kx] = fmaf(g1, p1[ky * Wp + kx], dk1[ky * 3 + kx]);
        dk2[ky * 3 + kx] = fmaf(g2, p2[ky * Wp + kx], dk2[ky * 3 + kx]);
      }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float a1 = block_sum(dk1[i], red);
    const float a2 = block_sum(dk2[i], red);
    if (threadIdx.x == 0) {
      dw_part[((size_t)b * HID2 + c) * 9 + i] = a1;
      dw_part[((size_t)b * HID2 + c + HID) * 9 + i] = a2;
    }
  }
}

// ------------------------------------------------------------------ 4-wide variants of the plane kernels (W % 4 == 0)
// Same arithmetic per output as the kernels above; one thread = four consecutive pixels of a row.  The padded plane has
// pitch W + 8 with the interior at column 4, so every row is 16-byte aligned: 128-bit global loads / stores, and a
// 3x3 window row for four outputs is one LDS.128 + two LDS.32 (2.25 shared loads per output and tap row instead of 9,
// one integer division per four outputs).  r1: gdfn_gate_fwd 92 us, dwconv 57 us at B = 16, 3-5x off their HBM floors.
__device__ __forceinline__ void load_plane4(float* s, const float* __restrict__ g, int H, int W) {
  const int PW = W + 8, W4 = W >> 2;
  for (int i = threadIdx.x; i < PW; i += blockDim.x) { s[i] = 0.f; s[(H + 1) * PW + i] = 0.f; }
  for (int i = threadIdx.x; i < H; i += blockDim.x) { s[(i + 1) * PW + 3] = 0.f; s[(i + 1) * PW + 4 + W] = 0.f; }
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int i = threadIdx.x; i < H * W4; i += blockDim.x) {
    const int y = i / W4, xq = i - y * W4;
    *reinterpret_cast<float4*>(s + (y + 1) * PW + 4 + 4 * xq) = __ldg(g4 + i);
  }
}
__device__ __forceinline__ void zero_plane4(float* s, int H, int W) {
  const int tot4 = ((H + 2) * (W + 8)) >> 2;
  for (int i = threadIdx.x; i < tot4; i += blockDim.x) reinterpret_cast<float4*>(s)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
// the 3 x 6 input window of the four outputs (y, x0 .. x0 + 3)
__device__ __forceinline__ void window4(const float* s, int PW, int y, int x0, float (&v)[3][6]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float* p = s + (y + r) * PW + 4 + x0;
    const float4 m = *reinterpret_cast<const float4*>(p);
    v[r][0] = p[-1]; v[r][1] = m.x; v[r][2] = m.y; v[r][3] = m.z; v[r][4] = m.w; v[r][5] = p[4];
  }
}
__device__ __forceinline__ void conv9x4(const float (&v)[3][6], const float* k, float (&o)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = v[0][j] * k[0] + v[0][j + 1] * k[1] + v[0][j + 2] * k[2] + v[1][j] * k[3] + v[1][j + 1] * k[4] + v[1][j + 2] * k[5] +
           v[2][j] * k[6] + v[2][j + 1] * k[7] + v[2][j + 2] * k[8];
}
__device__ __forceinline__ void conv9x4_flipped(const float (&v)[3][6], const float* k, float (&o)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = v[0][j] * k[8] + v[0][j + 1] * k[7] + v[0][j + 2] * k[6] + v[1][j] * k[5] + v[1][j + 1] * k[4] + v[1][j + 2] * k[3] +
           v[2][j] * k[2] + v[2][j + 1] * k[1] + v[2][j + 2] * k[0];
}

// N block-wide sums with one barrier pair: warp shuffles, one smem slot per (warp, value), then thread i < N adds the
// warp partials in a fixed order.  Returns the total of value `threadIdx.x` to threads < N (0 elsewhere).
template <int N>
__device__ __forceinline__ float block_sums(const float (&v)[N], float* red /* [8][N] */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float t = warp_sum(v[i]);
    if (lane == 0) red[warp * N + i] = t;
  }
  __syncthreads();
  float tot = 0.f;
  if (threadIdx.x < N)
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) tot += red[wi * N + threadIdx.x];
  return tot;
}

__global__ void __launch_bounds__(256)
dwconv_fwd_v4_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out, float* __restrict__ out2,
                     float* __restrict__ sumsq, int C, int split, int H, int W, int sq_C) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[8];
  const int c = blockIdx.x, b = blockIdx.y, N = H * W, PW = W + 8, W4 = W >> 2;
  load_plane4(sm, in + ((size_t)b * C + c) * N, H, W);
  float k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
  __syncthreads();
  float* o = (c < split) ? out + ((size_t)b * split + c) * N : out2 + ((size_t)b * (C - split) + (c - split)) * N;
  float ss = 0.f;
  for (int i = threadIdx.x; i < H * W4; i += blockDim.x) {
    const int y = i / W4, x0 = 4 * (i - y * W4);
    float v[3][6], r[4];
    window4(sm, PW, y, x0, v);
    conv9x4(v, k, r);
    reinterpret_cast<float4*>(o)[i] = make_float4(r[0], r[1], r[2], r[3]);
    ss = fmaf(r[0], r[0], fmaf(r[1], r[1], fmaf(r[2], r[2], fmaf(r[3], r[3], ss))));
  }
  if (sumsq != nullptr && c < sq_C) {                     // block-uniform: |row|^2 of the first sq_C channels, [B][sq_C]
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) sumsq[(size_t)b * sq_C + c] = ss;
  }
}

__global__ void __launch_bounds__(256)
dwconv_bwd_v4_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ dout,
                     const float* __restrict__ dout2, float* __restrict__ din, float* __restrict__ dw_part, int C, int split,
                     int H, int W) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red9[8 * 9];
  const int c = blockIdx.x, b = blockIdx.y, N = H * W, PW = W + 8, P = (H + 2) * PW, W4 = W >> 2;
  float* s_in = sm;
  float* s_do = sm + P;
  load_plane4(s_in, in + ((size_t)b * C + c) * N, H, W);
  const float* dg = (c < split) ? dout + ((size_t)b * split + c) * N : dout2 + ((size_t)b * (C - split) + (c - split)) * N;
  load_plane4(s_do, dg, H, W);
  float k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
  __syncthreads();
  float dk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) dk[i] = 0.f;
  float* di = din + ((size_t)b * C + c) * N;
  for (int i = threadIdx.x; i < H * W4; i += blockDim.x) {
    const int y = i / W4, x0 = 4 * (i - y * W4);
    float vd[3][6], vi[3][6], r[4];
    window4(s_do, PW, y, x0, vd);
    conv9x4_flipped(vd, k, r);
    reinterpret_cast<float4*>(di)[i] = make_float4(r[0], r[1], r[2], r[3]);
    window4(s_in, PW, y, x0, vi);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float g = vd[1][j + 1];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) dk[ky * 3 + kx] = fmaf(g, vi[ky][j + kx], dk[ky * 3 + kx]);
    }
  }
  const float t = block_sums<9>(dk, red9);
  if (threadIdx.x < 9) dw_part[((size_t)b * C + c) * 9 + threadIdx.x] = t;
}

// gs != NULL: the gate is written as the bf16 hi | lo B operand of the project_out GEMM ([B][HID][2 * Np], zero padded to Np)
// instead of fp32 g -- no fp32 round trip, no split pass (inference / tensor-core path)
__global__ void __launch_bounds__(256)
gdfn_gate_fwd_v4_kernel(const float* __restrict__ tpre, const float* __restrict__ w, float* __restrict__ g,
                        __nv_bfloat16* __restrict__ gs, int Np, int H, int W) {
  extern __shared__ __align__(16) float sm[];
  const int c = blockIdx.x, b = blockIdx.y, N = H * W, PW = W + 8, P = (H + 2) * PW, W4 = W >> 2;
  load_plane4(sm, tpre + ((size_t)b * HID2 + c) * N, H, W);
  load_plane4(sm + P, tpre + ((size_t)b * HID2 + c + HID) * N, H, W);
  float k1[9], k2[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { k1[i] = __ldg(w + c * 9 + i); k2[i] = __ldg(w + (c + HID) * 9 + i); }
  __syncthreads();
  float4* o = reinterpret_cast<float4*>(g + ((size_t)b * HID + c) * N);
  for (int i = threadIdx.x; i < H * W4; i += blockDim.x) {
    const int y = i / W4, x0 = 4 * (i - y * W4);
    float v[3][6], t1[4], t2[4];
    window4(sm, PW, y, x0, v);
    conv9x4(v, k1, t1);
    window4(sm + P, PW, y, x0, v);
    conv9x4(v, k2, t2);
    const float4 r = make_float4(gelu_erf(t1[0]) * t2[0], gelu_erf(t1[1]) * t2[1], gelu_erf(t1[2]) * t2[2], gelu_erf(t1[3]) * t2[3]);
    if (gs == nullptr) {
      o[i] = r;
    } else {
      __nv_bfloat16* d = gs + ((size_t)b * HID + c) * 2 * Np + 4 * i;
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(r.x, r.y), h1 = __floats2bfloat162_rn(r.z, r.w);
      const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
      const __nv_bfloat162 l0 = __floats2bfloat162_rn(r.x - __uint_as_float(u0 << 16), r.y - __uint_as_float(u0 & 0xffff0000u));
      const __nv_bfloat162 l1 = __floats2bfloat162_rn(r.z - __uint_as_float(u1 << 16), r.w - __uint_as_float(u1 & 0xffff0000u));
      *reinterpret_cast<uint2*>(d) = make_uint2(u0, u1);
      *reinterpret_cast<uint2*>(d + Np) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    }
  }
  if (gs != nullptr) {                                             // the K-padding columns [N, Np) of both halves read as zero
    __nv_bfloat16* d = gs + ((size_t)b * HID + c) * 2 * Np;
    for (int i = N + threadIdx.x; i < Np; i += blockDim.x) { d[i] = __float2bfloat16(0.f); d[Np + i] = __float2bfloat16(0.f); }
  }
}

__global__ void __launch_bounds__(256)
gdfn_gate_bwd_v4_kernel(const float* __restrict__ tpre, const float* __restrict__ w, const float* __restrict__ dg,
                        float* __restrict__ dtpre, float* __restrict__ g_out, float* __restrict__ dw_part, int H, int W) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red18[8 * 18];
  const int c = blockIdx.x, b = blockIdx.y, N = H * W, PW = W + 8, P = (H + 2) * PW, W4 = W >> 2;
  float* s1 = sm;            // t_pre[c]
  float* s2 = sm + P;        // t_pre[c + HID]
  float* d1 = sm + 2 * P;    // dt1 (padded)
  float* d2 = sm + 3 * P;    // dt2
  load_plane4(s1, tpre + ((size_t)b * HID2 + c) * N, H, W);
  load_plane4(s2, tpre + ((size_t)b * HID2 + c + HID) * N, H, W);
  zero_plane4(d1, H, W);
  zero_plane4(d2, H, W);
  float k1[9], k2[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { k1[i] = __ldg(w + c * 9 + i); k2[i] = __ldg(w + (c + HID) * 9 + i); }
  __syncthreads();
  const float4* dgp = reinterpret_cast<const float4*>(dg + ((size_t)b * HID + c) * N);
  float4* go = reinterpret_cast<float4*>(g_out + ((size_t)b * HID + c) * N);
  for (int i = threadIdx.x; i < H * W4; i += blockDim.x) {
    const int y = i / W4, x0 = 4 * (i - y * W4);
    float v[3][6], t1[4], t2[4];
    window4(s1, PW, y, x0, v);
    conv9x4(v, k1, t1);
    window4(s2, PW, y, x0, v);
    conv9x4(v, k2, t2);
    const float4 d4 = __ldg(dgp + i);
    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
    float gg[4], a1[4], a2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      gg[j] = gelu_erf(t1[j]);
      a1[j] = d[j] * t2[j] * gelu_erf_grad(t1[j]);
      a2[j] = d[j] * gg[j];
    }
    go[i] = make_float4(gg[0] * t2[0], gg[1] * t2[1], gg[2] * t2[2], gg[3] * t2[3]);
    *reinterpret_cast<float4*>(d1 + (y + 1) * PW + 4 + x0) = make_float4(a1[0], a1[1], a1[2], a1[3]);
    *reinterpret_cast<float4*>(d2 + (y + 1) * PW + 4 + x0) = make_float4(a2[0], a2[1], a2[2], a2[3]);
  }
  __syncthreads();
  float dkk[18];             // [0,9) channel c, [9,18) channel c + HID
#pragma unroll
  for (int i = 0; i < 18; ++i) dkk[i] = 0.f;
  float4* o1 = reinterpret_cast<float4*>(dtpre + ((size_t)b * HID2 + c) * N);
  float4* o2 = reinterpret_cast<float4*>(dtpre + ((size_t)b * HID2 + c + HID) * N);
  for (int i = threadIdx.x; i < H * W4; i += blockDim.x) {
    const int y = i / W4, x0 = 4 * (i - y * W4);
    float vd[3][6], vi[3][6], r[4];
    window4(d1, PW, y, x0, vd);
    conv9x4_flipped(vd, k1, r);
    o1[i] = make_float4(r[0], r[1], r[2], r[3]);
    window4(s1, PW, y, x0, vi);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gq = vd[1][j + 1];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) dkk[ky * 3 + kx] = fmaf(gq, vi[ky][j + kx], dkk[ky * 3 + kx]);
    }
    window4(d2, PW, y, x0, vd);
    conv9x4_flipped(vd, k2, r);
    o2[i] = make_float4(r[0], r[1], r[2], r[3]);
    window4(s2, PW, y, x0, vi);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gq = vd[1][j + 1];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) dkk[9 + ky * 3 + kx] = fmaf(gq, vi[ky][j + kx], dkk[9 + ky * 3 + kx]);
    }
  }
  const float tot = block_sums<18>(dkk, red18);
  if (threadIdx.x < 9) dw_part[((size_t)b * HID2 + c) * 9 + threadIdx.x] = tot;
  else if (threadIdx.x < 18) dw_part[((size_t)b * HID2 + c + HID) * 9 + threadIdx.x - 9] = tot;
}

// ------------------------------------------------------------------ MDTA channel attention, per (sample, head)
// forward: attn = softmax_d( G[c][d] / (|q_c| |k_d|) * temperature )   (PromptInteract.py:421-425), then folds
// it into project_out:  M[b][o][h*64+d] = sum_c Wo[o][h*64+c] attn[c][d], so that  z = M[b] v.
// Grid (B * HEADS, 4): every CTA redoes the (tiny) softmax, CTA y folds it into output rows o in [32 y, 32 y + 32);
// one thread = one column d x eight rows o, weights read as warp-uniform 128-bit loads (r1: one CTA per (sample, head)
// with a 64-deep dependent global-load loop per output took 43 us on 32 SMs).
__global__ void __launch_bounds__(256)
mdta_attn_fwd_kernel(const float* __restrict__ G, const float* __restrict__ sq, const float* __restrict__ sk,
                     const float* __restrict__ temperature, const float* __restrict__ wo, float* __restrict__ attn,
                     float* __restrict__ M) {
  __shared__ float sa[HD][HD + 1];
  const int bh = blockIdx.x, b = bh / HEADS, h = bh % HEADS;
  const float tau = __ldg(temperature + h);
  const float* g = G + (size_t)bh * HD * HD;
  // one warp per row c (8 warps -> 8 rows per pass); lane handles d and d + 32
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < HD; c += 8) {
    const float nq = fmaxf(sqrtf(__ldg(sq + (size_t)b * DIM + h * HD + c)), NORM_EPS);
    float s0 = __ldg(g + c * HD + lane) / (nq * fmaxf(sqrtf(__ldg(sk + (size_t)b * DIM + h * HD + lane)), NORM_EPS)) * tau;
    float s1 = __ldg(g + c * HD + lane + 32) /
               (nq * fmaxf(sqrtf(__ldg(sk + (size_t)b * DIM + h * HD + lane + 32)), NORM_EPS)) * tau;
    const float m = warp_max(fmaxf(s0, s1));
    s0 = expf(s0 - m);
    s1 = expf(s1 - m);
    const float l = warp_sum(s0 + s1);
    s0 /= l;
    s1 /= l;
    sa[c][lane] = s0;
    sa[c][lane + 32] = s1;
    if (blockIdx.y == 0) {
      attn[(size_t)bh * HD * HD + c * HD + lane] = s0;
      attn[(size_t)bh * HD * HD + c * HD + lane + 32] = s1;
    }
  }
  __syncthreads();
  const int d = threadIdx.x & 63, o0 = 32 * blockIdx.y + 8 * (threadIdx.x >> 6);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 2
  for (int c = 0; c < HD; c += 4) {
    const float a0 = sa[c][d], a1 = sa[c + 1][d], a2 = sa[c + 2][d], a3 = sa[c + 3][d];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(wo + (size_t)(o0 + k) * DIM + h * HD + c));
      acc[k] = fmaf(w4.x, a0, fmaf(w4.y, a1, fmaf(w4.z, a2, fmaf(w4.w, a3, acc[k]))));
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) M[(size_t)b * DIM * DIM + (size_t)(o0 + k) * DIM + h * HD + d] = acc[k];
}

// backward: P[b][o][j] = sum_n dz[b][o][n] v[b][j][n]  (dz = gradient at project_out's output)
//   dattn[c][d] = sum_o Wo[o][h64+c] P[o][h64+d]          dWo_part[b][o][h64+c] = sum_d P[o][h64+d] attn[c][d]
//   dS = attn o (dattn - rowsum(dattn o attn))
//   S = tau G / (nq nk)  =>  dG' = dS tau/(nq nk) ; dtau_part = sum dS G/(nq nk) ;
//   d nq_c = -sum_d dS S / nq_c ; d nk_d = -sum_c dS S / nk_d ; a_q = d nq / |q| (0 when |q| <= eps), same for k
__global__ void __launch_bounds__(256)
mdta_attn_bwd_kernel(const float* __restrict__ P, const float* __restrict__ attn, const float* __restrict__ G,
                     const float* __restrict__ sq, const float* __restrict__ sk, const float* __restrict__ temperature,
                     const float* __restrict__ wo, float* __restrict__ dGs, float* __restrict__ aq, float* __restrict__ ak,
                     float* __restrict__ dtau_part, float* __restrict__ dwo_part) {
  __shared__ float sa[HD][HD + 1];    // attn
  __shared__ float sd[HD][HD + 1];    // dattn, then dS
  __shared__ float snq[HD], snk[HD];  // max(|q_c|, eps), max(|k_d|, eps)
  __shared__ float red[8];
  const int bh = blockIdx.x, b = bh / HEADS, h = bh % HEADS;
  const float tau = __ldg(temperature + h);
  const float* p = P + (size_t)b * DIM * DIM;
  const float* Gb = G + (size_t)bh * HD * HD;
  if (threadIdx.x < HD) snq[threadIdx.x] = fmaxf(sqrtf(__ldg(sq + (size_t)b * DIM + h * HD + threadIdx.x)), NORM_EPS);
  else if (threadIdx.x < 2 * HD) snk[threadIdx.x - HD] = fmaxf(sqrtf(__ldg(sk + (size_t)b * DIM + h * HD + threadIdx.x - HD)), NORM_EPS);
  __syncthreads();
  auto sS = [&](int c, int d) { return __ldg(Gb + c * HD + d) / (snq[c] * snk[d]); };   // S / tau
  {
    float4 t[4];                                         // all loads in flight before the first use
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = __ldg(reinterpret_cast<const float4*>(attn + (size_t)bh * HD * HD) + threadIdx.x + 256 * j);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = 4 * (threadIdx.x + 256 * j), c = i / HD, d = i % HD;
      sa[c][d] = t[j].x; sa[c][d + 1] = t[j].y; sa[c][d + 2] = t[j].z; sa[c][d + 3] = t[j].w;
    }
  }
  if (blockIdx.y > 0) {
    // CTAs 1..4: the project_out weight-gradient partial for output rows o in [32 (y - 1), 32 y): one thread = one
    // column c x eight rows o, P read as warp-uniform 128-bit loads
    __syncthreads();
    const int c = threadIdx.x & 63, o0 = 32 * (blockIdx.y - 1) + 8 * (threadIdx.x >> 6);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 2
    for (int d = 0; d < HD; d += 4) {
      const float a0 = sa[c][d], a1 = sa[c][d + 1], a2 = sa[c][d + 2], a3 = sa[c][d + 3];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(o0 + k) * DIM + h * HD + d));
        acc[k] = fmaf(p4.x, a0, fmaf(p4.y, a1, fmaf(p4.z, a2, fmaf(p4.w, a3, acc[k]))));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) dwo_part[(size_t)b * DIM * DIM + (size_t)(o0 + k) * DIM + h * HD + c] = acc[k];
    return;
  }
  {
    // CTA 0: dattn[c][d] = sum_o Wo[o][h64+c] P[o][h64+d]: one thread = one column d x sixteen rows c.  The two
    // [128][64] operand slices are staged in shared memory with every load in flight at once (a 128-deep loop of
    // dependent L2 round trips was 40 % of this kernel's 84 us)
    extern __shared__ __align__(16) float dyn[];
    float4* swo4 = reinterpret_cast<float4*>(dyn);
    float4* sp4 = swo4 + DIM * HD / 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = threadIdx.x + 256 * j, o = i >> 4, q = i & 15;
      swo4[i] = __ldg(reinterpret_cast<const float4*>(wo + (size_t)o * DIM + h * HD) + q);
      sp4[i] = __ldg(reinterpret_cast<const float4*>(p + (size_t)o * DIM + h * HD) + q);
    }
    __syncthreads();
    const float* sp = reinterpret_cast<const float*>(sp4);
    const int d = threadIdx.x & 63, c0 = 16 * (threadIdx.x >> 6);
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
#pragma unroll 4
    for (int o = 0; o < DIM; ++o) {
      const float pv = sp[o * HD + d];
      const float4* w4 = swo4 + (o * HD + c0) / 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w = w4[q];
        acc[4 * q] = fmaf(w.x, pv, acc[4 * q]);
        acc[4 * q + 1] = fmaf(w.y, pv, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(w.z, pv, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(w.w, pv, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) sd[c0 + k][d] = acc[k];
  }
  __syncthreads();
  // softmax backward, one warp per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < HD; c += 8) {
    const float r = warp_sum(sd[c][lane] * sa[c][lane] + sd[c][lane + 32] * sa[c][lane + 32]);
    sd[c][lane] = sa[c][lane] * (sd[c][lane] - r);
    sd[c][lane + 32] = sa[c][lane + 32] * (sd[c][lane + 32] - r);
  }
  __syncthreads();
  float tsum = 0.f;
  for (int i = threadIdx.x; i < HD * HD; i += 256) {
    const int c = i / HD, d = i % HD;
    tsum = fmaf(sd[c][d], sS(c, d), tsum);
    dGs[(size_t)bh * HD * HD + i] = sd[c][d] * tau / (snq[c] * snk[d]);
  }
  tsum = block_sum(tsum, red);
  if (threadIdx.x == 0) dtau_part[bh] = tsum;
  if (threadIdx.x < HD) {
    const int c = threadIdx.x;
    float s = 0.f;
    for (int d = 0; d < HD; ++d) s = fmaf(sd[c][d], sS(c, d), s);
    const float n2 = __ldg(sq + (size_t)b * DIM + h * HD + c), nq = sqrtf(n2);
    // d|q_c| = -tau * s / nq ;  dq += d|q| * q / |q|
    aq[(size_t)b * DIM + h * HD + c] = (nq > NORM_EPS) ? (-tau * s / (nq * nq)) : 0.f;
  } else if (threadIdx.x < 2 * HD) {
    const int d = threadIdx.x - HD;
    float s = 0.f;
    for (int c = 0; c < HD; ++c) s = fmaf(sd[c][d], sS(c, d), s);
    const float n2 = __ldg(sk + (size_t)b * DIM + h * HD + d), nk = sqrtf(n2);
    ak[(size_t)b * DIM + h * HD + d] = (nk > NORM_EPS) ? (-tau * s / (nk * nk)) : 0.f;
  }
}

// y[b][c][n] += a[b][c] * x[b][c][n]
__global__ void row_axpy_kernel(const float* __restrict__ a, const float* __restrict__ x, float* __restrict__ y, int N) {
  const size_t row = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) y[row * N + n] = fmaf(__ldg(a + row), __ldg(x + row * N + n), y[row * N + n]);
}

// ------------------------------------------------------------------ LayerNorm backward
// dx = rstd (dn g - mean_c(dn g) - xhat mean_c(dn g xhat)) (+ add), dn = gradient at the LN output
__global__ void __launch_bounds__(LN_PX * LN_GR)
ln_bwd_dx_kernel(const float* __restrict__ dn, const float* __restrict__ x, const float* __restrict__ mean,
                 const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ add,
                 float* __restrict__ dx, int C, int N) {
  __shared__ float r1[LN_GR][LN_PX], r2[LN_GR][LN_PX];
  const int b = blockIdx.y, px = threadIdx.x & (LN_PX - 1), gr = threadIdx.x / LN_PX;
  const int n = blockIdx.x * LN_PX + px;
  const bool ok = n < N;
  const size_t base = (size_t)b * C * N + (size_t)gr * LN_CPT * N + (ok ? n : 0);
  const float mu = ok ? __ldg(mean + (size_t)b * N + n) : 0.f, rs = ok ? __ldg(rstd + (size_t)b * N + n) : 0.f;
  float g[LN_CPT], xh[LN_CPT];
#pragma unroll
  for (int j = 0; j < LN_CPT; ++j) {
    g[j] = __ldg(dn + base + (size_t)j * N) * __ldg(gamma + gr * LN_CPT + j);
    xh[j] = (__ldg(x + base + (size_t)j * N) - mu) * rs;
  }
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < LN_CPT; ++j) {
    s1 += g[j];
    s2 = fmaf(g[j], xh[j], s2);
  }
  r1[gr][px] = s1;
  r2[gr][px] = s2;
  __syncthreads();
  s1 = 0.f;
  s2 = 0.f;
#pragma unroll
  for (int k = 0; k < LN_GR; ++k) { s1 += r1[k][px]; s2 += r2[k][px]; }
  s1 /= (float)C;
  s2 /= (float)C;
  if (!ok) return;
#pragma unroll
  for (int j = 0; j < LN_CPT; ++j) {
    float v = rs * (g[j] - s1 - xh[j] * s2);
    if (add != nullptr) v += __ldg(add + base + (size_t)j * N);
    dx[base + (size_t)j * N] = v;
  }
}
// per (b, c): dgamma_part = sum_n dn xhat, dbeta_part = sum_n dn
__global__ void __launch_bounds__(256)
ln_bwd_param_kernel(const float* __restrict__ dn, const float* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ rstd, float* __restrict__ dg_part, float* __restrict__ db_part, int C, int N) {
  __shared__ float red[8];
  const int c = blockIdx.x, b = blockIdx.y;
  const size_t row = ((size_t)b * C + c) * N;
  float sg = 0.f, sb = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float d = __ldg(dn + row + n);
    const float xh = (__ldg(x + row + n) - __ldg(mean + (size_t)b * N + n)) * __ldg(rstd + (size_t)b * N + n);
    sg = fmaf(d, xh, sg);
    sb += d;
  }
  sg = block_sum(sg, red);
  sb = block_sum(sb, red);
  if (threadIdx.x == 0) { dg_part[(size_t)b * C + c] = sg; db_part[(size_t)b * C + c] = sb; }
}

// ------------------------------------------------------------------ host-side launch helpers
int launch_ln_stats(const float* x, float* mean, float* rstd, int B, int C, int N, cudaStream_t st) {
  if (C != DIM) { emip_set_error("ln_stats: C=%d unsupported", C); return EMIP_ENOSYS; }
  ln_stats_kernel<<<dim3((N + LN_PX - 1) / LN_PX, B), LN_PX * LN_GR, 0, st>>>(x, mean, rstd, C, N);
  EMIP_CHECK_LAUNCH("ln_stats");
  return EMIP_OK;
}
// the statistics of two tensors of one shape (norm1 of x, norm2 of x1) in one launch
int launch_ln_stats2(const float* x, float* mean, float* rstd, const float* x2, float* mean2, float* rstd2, int B, int C, int N,
                     cudaStream_t st) {
  if (C != DIM) { emip_set_error("ln_stats: C=%d unsupported", C); return EMIP_ENOSYS; }
  ln_stats_kernel<<<dim3((N + LN_PX - 1) / LN_PX, B, 2), LN_PX * LN_GR, 0, st>>>(x, mean, rstd, C, N, x2, mean2, rstd2);
  EMIP_CHECK_LAUNCH("ln_stats");
  return EMIP_OK;
}
// W % 4 == 0 selects the 4-wide plane kernels (pitch W + 8), anything else the scalar ones (pitch W + 2)
size_t plane_smem(int H, int W, int planes) { return sizeof(float) * (size_t)planes * (H + 2) * (W + ((W & 3) ? 2 : 8)); }

template <typename K>
int ensure_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) EMIP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return EMIP_OK;
}

// carve helper
struct Carver {
  char* p;
  size_t left;
  bool ok;
  float* take(size_t n_floats) {
    size_t bytes = emip_align_up(n_floats * sizeof(float), 256);
    if (bytes > left) { ok = false; return nullptr; }
    float* r = reinterpret_cast<float*>(p);
    p += bytes;
    left -= bytes;
    return r;
  }
};
size_t al(size_t n_floats) { return emip_align_up(n_floats * sizeof(float), 256); }

// parameter order = the order of the reference's state_dict (include/emip_b200.h)
enum {
  P_N1W, P_N1B, P_N2W, P_N2B, P_N3W, P_N3B, P_TEMP, P_QW, P_QDW, P_KVW, P_KVDW, P_POW, P_FIW, P_FDW, P_FOW, P_COUNT
};

struct Saved {          // activations kept for the backward pass (caller-owned buffer)
  float *mean1, *rstd1, *mean2, *rstd2, *mean3, *rstd3;
  float *qpre, *kvpre, *q, *k, *v, *sq, *sk, *G, *attn, *M, *y, *tpre;
};
size_t saved_bytes(int B, int N) {
  size_t s = 6 * al((size_t)B * N);
  s += al((size_t)B * DIM * N) * 5 + al((size_t)B * 2 * DIM * N);          // qpre q k v y + kvpre
  s += 2 * al((size_t)B * DIM) + 2 * al((size_t)B * HEADS * HD * HD) + al((size_t)B * DIM * DIM);
  s += al((size_t)B * HID2 * N);
  return s;
}
bool carve_saved(void* buf, size_t bytes, int B, int N, Saved* s) {
  Carver c{static_cast<char*>(buf), bytes, true};
  s->mean1 = c.take((size_t)B * N); s->rstd1 = c.take((size_t)B * N);
  s->mean2 = c.take((size_t)B * N); s->rstd2 = c.take((size_t)B * N);
  s->mean3 = c.take((size_t)B * N); s->rstd3 = c.take((size_t)B * N);
  s->qpre = c.take((size_t)B * DIM * N); s->kvpre = c.take((size_t)B * 2 * DIM * N);
  s->q = c.take((size_t)B * DIM * N); s->k = c.take((size_t)B * DIM * N); s->v = c.take((size_t)B * DIM * N);
  s->sq = c.take((size_t)B * DIM); s->sk = c.take((size_t)B * DIM);
  s->G = c.take((size_t)B * HEADS * HD * HD); s->attn = c.take((size_t)B * HEADS * HD * HD);
  s->M = c.take((size_t)B * DIM * DIM);
  s->y = c.take((size_t)B * DIM * N); s->tpre = c.take((size_t)B * HID2 * N);
  return c.ok;
}
// pixel-axis splits of the weight-gradient GEMMs: enough (batch x split x output tiles) CTAs for two per SM
int nt_split_simt(int B, int M, int K) {
  const int tiles = ((M + 127) / 128) * ((K + 127) / 128);
  int s = (2 * emip_num_sms() + B * tiles - 1) / (B * tiles);
  return s < 1 ? 1 : (s > 16 ? 16 : s);
}
size_t nt_part_floats(int B) {       // largest partial buffer any of the backward's NT GEMMs needs (CUDA-core path)
  size_t m = 0;
  const int shp[5][2] = {{DIM, HID}, {HID2, DIM}, {DIM, DIM}, {2 * DIM, DIM}, {DIM, DIM}};
  for (auto& sh : shp) {
    const size_t v = (size_t)B * nt_split_simt(B, sh[0], sh[1]) * sh[0] * sh[1];
    m = v > m ? v : m;
  }
  return m;
}
constexpr int G_SPLIT = 8;      // pixel-axis splits of the 64 x 64 Gram matrices

int check_common(const char* who, int B, int H, int W) {
  EMIP_CHECK_ARG(B >= 0 && H > 0 && W > 0, "%s: bad shape B=%d H=%d W=%d", who, B, H, W);
  if (plane_smem(H, W, 4) > 200 * 1024) {
    emip_set_error("%s: %dx%d feature maps do not fit the plane-resident depthwise kernels", who, H, W);
    return EMIP_ENOSYS;
  }
  return EMIP_OK;
}

}  // namespace

// ------------------------------------------------------------------ public GEMM helpers (gemm_simt.cuh)
// Set by emip_injector_fwd_ex for the duration of a forward call on this thread: scratch for the tensor-core GEMMs.
struct TcScope {
  void* ws = nullptr;
  size_t bytes = 0;
};
static thread_local TcScope g_tc;

int gemm_nn(const GemmNN& a, cudaStream_t st) {
  if (a.B == 0 || a.M == 0 || a.N == 0) return EMIP_OK;
  if (g_tc.ws != nullptr && gemm_nn_tc_supported(a) &&
      g_tc.bytes >= gemm_nn_tc_scratch_bytes(a.B, a.M, a.K, a.N, a.w_stride_b != 0))
    return gemm_nn_tc(a, g_tc.ws, g_tc.bytes, st);
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  const bool fast = a.N % 4 == 0 && a.ldx % 4 == 0 && a.ldy % 4 == 0 && a.x_stride_b % 4 == 0 && a.y_stride_b % 4 == 0 &&
                    al16(a.x) && al16(a.y) && al16(a.w) && a.w_stride_b % 4 == 0 &&
                    (a.w_trans ? a.ldw % 4 == 0 : a.ldw % 4 == 0) &&
                    (!a.res || (a.ldr % 4 == 0 && a.res_stride_b % 4 == 0 && al16(a.res))) &&
                    (!a.mean || (al16(a.mean) && al16(a.rstd))) && a.M >= 128 && a.M % 4 == 0;
  if (fast) {
    dim3 grid((a.N + FB - 1) / FB, (a.M + FB - 1) / FB, a.B);
    gemm_nn_fast_kernel<<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid((a.N + GB - 1) / GB, (a.M + GB - 1) / GB, a.B);
    gemm_nn_kernel<<<grid, 256, 0, st>>>(a);
  }
  EMIP_CHECK_LAUNCH("gemm_nn");
  return EMIP_OK;
}
// nsplit partial results per batch entry: c must hold B*nsplit matrices (c_stride_b apart)
// pixel-axis splits of a weight-gradient GEMM: the tensor-core path produces one partial per batch entry
// tensor-core path: enough K splits to give every SM a tile (r2d: one partial per sample left the weight-gradient
// GEMMs on 16-96 CTAs for 27-30 us each); never more than the CUDA-core path's count, which sizes the partial buffer
static int nt_split(int B, int M, int K) {
  const int simt = nt_split_simt(B, M, K);
  if (g_tc.ws == nullptr) return simt;
  const int tiles = B * ((M + 127) / 128) * ((K + 127) / 128);
  int s = emip_num_sms() / tiles;
  s = s < 1 ? 1 : (s > 8 ? 8 : s);
  return s < simt ? s : simt;
}

static int gemm_nt_split(const GemmNT& a, int nsplit, cudaStream_t st) {
  if (a.B == 0 || a.M == 0 || a.K == 0) return EMIP_OK;
  if (g_tc.ws != nullptr && gemm_nt_tc_supported(a) && g_tc.bytes >= gemm_nt_tc_scratch_bytes(a.B, a.M, a.K, a.N))
    return gemm_nt_tc(a, g_tc.ws, g_tc.bytes, st, nsplit);
  GemmNTk p;
  p.g = a;
  p.nsplit = nsplit;
  p.chunk = ((a.N + nsplit - 1) / nsplit + 31) / 32 * 32;
  auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  const bool fast = a.N % 4 == 0 && a.lda % 4 == 0 && a.ldb % 4 == 0 && a.a_stride_b % 4 == 0 && a.b_stride_b % 4 == 0 &&
                    al16(a.a) && al16(a.bm) && (!a.mean || (al16(a.mean) && al16(a.rstd))) && a.M >= 128 && a.K >= 128;
  if (fast) {
    dim3 grid((a.K + FB - 1) / FB, (a.M + FB - 1) / FB, a.B * nsplit);
    gemm_nt_fast_kernel<<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((a.K + GB - 1) / GB, (a.M + GB - 1) / GB, a.B * nsplit);
    gemm_nt_kernel<<<grid, 256, 0, st>>>(p);
  }
  EMIP_CHECK_LAUNCH("gemm_nt");
  return EMIP_OK;
}
int gemm_nt(const GemmNT& a, cudaStream_t st) { return gemm_nt_split(a, 1, st); }
int reduce_batch(const float* in, long long stride, float* out, int B, long long n, int accumulate, cudaStream_t st) {
  if (n == 0) return EMIP_OK;
  reduce_batch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, stride, out, B, n, accumulate);
  EMIP_CHECK_LAUNCH("reduce_batch");
  return EMIP_OK;
}

// largest scratch any of the forward / backward tensor-core GEMMs needs (they run one after the other)
static size_t tc_scratch_bytes(int B, int N) {
  size_t m = 0;
  // gemm_nn: batch factor, M, K, per-sample W
  const int nn[9][4] = {{1, DIM, DIM, 0}, {1, 2 * DIM, DIM, 0}, {1, DIM, DIM, 1}, {1, HID2, DIM, 0}, {1, DIM, HID, 0},
                        {1, HID, DIM, 0}, {1, DIM, HID2, 0}, {1, DIM, 2 * DIM, 0}, {HEADS, HD, HD, 1}};
  for (auto& sh : nn) {
    const size_t b = gemm_nn_tc_scratch_bytes(B * sh[0], sh[1], sh[2], N, sh[3] != 0);
    if (b > m) m = b;
  }
  // gemm_nt: M, K
  const int nt[5][2] = {{DIM, HID}, {HID2, DIM}, {DIM, DIM}, {2 * DIM, DIM}, {DIM, DIM}};
  for (auto& sh : nt) {
    const size_t b = gemm_nt_tc_scratch_bytes(B, sh[0], sh[1], N);
    if (b > m) m = b;
  }
  return emip_align_up(m, 256);
}

// ------------------------------------------------------------------ C ABI
extern "C" size_t emip_injector_saved_bytes(int B, int H, int W) {
  if (B < 0 || H <= 0 || W <= 0) return 0;
  return saved_bytes(B, H * W);
}
extern "C" size_t emip_injector_workspace(int B, int H, int W) {
  if (B < 0 || H <= 0 || W <= 0) return 0;
  const size_t N = (size_t)H * W;
  // forward: g [B,340,N]; backward: g, dg, dtpre [B,680,N], 4 x [B,128,N], dkvpre [B,256,N], dn [B,128,N], partials
  size_t s = 2 * al((size_t)B * HID * N) + al((size_t)B * HID2 * N) + 6 * al((size_t)B * DIM * N) +
             al((size_t)B * 2 * DIM * N);
  s += al(nt_part_floats(B));                                      // largest weight-gradient partial
  s += al((size_t)B * DIM * DIM) * 2 + al((size_t)B * HEADS * HD * HD) + 4 * al((size_t)B * DIM);
  s += al((size_t)B * HID2 * 9) + 2 * al((size_t)B * DIM) + al((size_t)B * HEADS);
  s += al((size_t)B * HEADS * G_SPLIT * HD * HD);
  s += tc_scratch_bytes(B, (int)N);                                // bf16 hi|lo operands of the tensor-core forward GEMMs
  return s;
}

extern "C" int emip_injector_fwd(const float* x, const float* x1, const float* const* params, float* out, void* saved,
                                 size_t saved_bytes_in, void* workspace, size_t ws_bytes, int B, int H, int W,
                                 void* stream) {
  return emip_injector_fwd_ex(x, x1, params, out, saved, saved_bytes_in, workspace, ws_bytes, B, H, W, 0, stream);
}

extern "C" int emip_injector_fwd_ex(const float* x, const float* x1, const float* const* params, float* out, void* saved,
                                    size_t saved_bytes_in, void* workspace, size_t ws_bytes, int B, int H, int W, int flags,
                                    void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && x1 && params && out && saved && workspace, "injector_fwd: null pointer");
  int rc = check_common("injector_fwd", B, H, W);
  if (rc) return rc;
  for (int i = 0; i < P_COUNT; ++i) EMIP_CHECK_ARG(params[i] != nullptr, "injector_fwd: parameter %d is NULL", i);
  cudaStream_t st = (cudaStream_t)stream;
  const int N = H * W;
  Saved s;
  if (saved_bytes_in < saved_bytes(B, N) || !carve_saved(saved, saved_bytes_in, B, N, &s)) {
    emip_set_error("injector_fwd: saved buffer too small (%zu < %zu)", saved_bytes_in, saved_bytes(B, N));
    return EMIP_ENOMEM;
  }
  if (ws_bytes < emip_injector_workspace(B, H, W)) {
    emip_set_error("injector_fwd: workspace too small");
    return EMIP_ENOMEM;
  }
  Carver w{static_cast<char*>(workspace), ws_bytes, true};
  float* g = w.take((size_t)B * HID * N);
  // The five 1x1 convolutions run on the tensor cores (3-term split-bf16, fp32 accumulate; gemm_tc.cuh) unless the
  // caller asks for the exact-fp32 CUDA-core GEMMs.  LayerNorm is applied while the activations are re-laid out.
  struct Scope {
    ~Scope() { g_tc = TcScope(); }
  } scope;
  if (!(flags & EMIP_FLAG_EXACT_FP32)) {
    g_tc.bytes = tc_scratch_bytes(B, N);
    g_tc.ws = reinterpret_cast<char*>(workspace) + (ws_bytes - g_tc.bytes) / 256 * 256;   // the tail of the workspace
  }

  // norm1 / norm2 statistics (PromptInteract.py:447, :346-349)
  if ((rc = launch_ln_stats2(x, s.mean1, s.rstd1, x1, s.mean2, s.rstd2, B, DIM, N, st))) return rc;
  // q = q_dwconv(q(LN1 x)), kv = kv_dwconv(kv(LN2 x1))      (:413-415)
  GemmNN a = {};
  a.B = B; a.N = N; a.K = DIM; a.ldx = N; a.x_stride_b = (long long)DIM * N; a.ldy = N;
  a.w = params[P_QW]; a.ldw = DIM; a.M = DIM; a.x = x; a.mean = s.mean1; a.rstd = s.rstd1;
  a.gamma = params[P_N1W]; a.beta = params[P_N1B]; a.y = s.qpre; a.y_stride_b = (long long)DIM * N;
  if ((rc = gemm_nn(a, st))) return rc;
  a.w = params[P_KVW]; a.M = 2 * DIM; a.x = x1; a.mean = s.mean2; a.rstd = s.rstd2;
  a.gamma = params[P_N2W]; a.beta = params[P_N2B]; a.y = s.kvpre; a.y_stride_b = (long long)2 * DIM * N;
  if ((rc = gemm_nn(a, st))) return rc;
  const size_t sm1 = plane_smem(H, W, 1);
  const bool v4 = (W & 3) == 0;
  if ((rc = v4 ? ensure_smem(dwconv_fwd_v4_kernel, sm1) : ensure_smem(dwconv_fwd_kernel, sm1))) return rc;
  if (v4) dwconv_fwd_v4_kernel<<<dim3(DIM, B), 256, sm1, st>>>(s.qpre, params[P_QDW], s.q, nullptr, s.sq, DIM, DIM, H, W, DIM);
  else dwconv_fwd_kernel<<<dim3(DIM, B), 256, sm1, st>>>(s.qpre, params[P_QDW], s.q, nullptr, s.sq, DIM, DIM, H, W, DIM);
  EMIP_CHECK_LAUNCH("dwconv q");
  // k = channels [0,128), v = [128,256) of kv (:415); only k is L2-normalised: |k_d|^2 over the pixels (F.normalize, :422) comes
  // out of this kernel like |q_c|^2 out of the one above (r5: a separate row_sumsq pass over k, 49 us at 64 pairs, is gone)
  if (v4) dwconv_fwd_v4_kernel<<<dim3(2 * DIM, B), 256, sm1, st>>>(s.kvpre, params[P_KVDW], s.k, s.v, s.sk, 2 * DIM, DIM, H, W, DIM);
  else dwconv_fwd_kernel<<<dim3(2 * DIM, B), 256, sm1, st>>>(s.kvpre, params[P_KVDW], s.k, s.v, s.sk, 2 * DIM, DIM, H, W, DIM);
  EMIP_CHECK_LAUNCH("dwconv kv");
  // G[c][d] = sum_n q[c][n] k[d][n] per (sample, head)      (:421-424, normalisation folded in afterwards)
  {
    GemmNT t = {};
    t.B = B * HEADS; t.M = HD; t.K = HD; t.N = N;
    t.a = s.q; t.a_stride_b = (long long)HD * N; t.lda = N;
    t.bm = s.k; t.b_stride_b = (long long)HD * N; t.ldb = N;
    // 2B problems of 64 x 64 x N: split the pixel axis so that the launch fills the chip, then sum the partials
    float* gpart = w.take((size_t)B * HEADS * G_SPLIT * HD * HD);
    if (gpart == nullptr) { emip_set_error("injector_fwd: workspace carve failed"); return EMIP_ENOMEM; }
    t.c = gpart; t.c_stride_b = HD * HD; t.ldc = HD;
    if ((rc = gemm_nt_split(t, G_SPLIT, st))) return rc;
    sum_splits_kernel<<<(B * HEADS * HD * HD + 255) / 256, 256, 0, st>>>(gpart, s.G, B * HEADS, G_SPLIT, HD * HD);
    EMIP_CHECK_LAUNCH("sum_splits");
  }
  mdta_attn_fwd_kernel<<<dim3(B * HEADS, 4), 256, 0, st>>>(s.G, s.sq, s.sk, params[P_TEMP], params[P_POW], s.attn, s.M);
  EMIP_CHECK_LAUNCH("mdta_attn_fwd");
  // y = x + project_out(attn @ v) = x + M[b] v             (:427-431, :447)
  a = {};
  a.B = B; a.M = DIM; a.K = DIM; a.N = N; a.w = s.M; a.w_stride_b = DIM * DIM; a.ldw = DIM;
  a.x = s.v; a.x_stride_b = (long long)DIM * N; a.ldx = N;
  a.res = x; a.res_stride_b = (long long)DIM * N; a.ldr = N;
  a.y = s.y; a.y_stride_b = (long long)DIM * N; a.ldy = N;
  if ((rc = gemm_nn(a, st))) return rc;
  // GDFN: out = y + project_out(gelu(t1) * t2), t = dwconv(project_in(LN3 y))     (:380-385, :448)
  if ((rc = launch_ln_stats(s.y, s.mean3, s.rstd3, B, DIM, N, st))) return rc;
  a = {};
  a.B = B; a.M = HID2; a.K = DIM; a.N = N; a.w = params[P_FIW]; a.ldw = DIM;
  a.x = s.y; a.x_stride_b = (long long)DIM * N; a.ldx = N; a.mean = s.mean3; a.rstd = s.rstd3;
  a.gamma = params[P_N3W]; a.beta = params[P_N3B];
  a.y = s.tpre; a.y_stride_b = (long long)HID2 * N; a.ldy = N;
  if ((rc = gemm_nn(a, st))) return rc;
  const size_t sm2 = plane_smem(H, W, 2);
  if ((rc = v4 ? ensure_smem(gdfn_gate_fwd_v4_kernel, sm2) : ensure_smem(gdfn_gate_fwd_kernel, sm2))) return rc;
  a = {};
  a.B = B; a.M = DIM; a.K = HID; a.N = N; a.w = params[P_FOW]; a.ldw = HID;
  a.x = g; a.x_stride_b = (long long)HID * N; a.ldx = N;
  // tensor-core path: the gate kernel writes the GEMM's bf16 hi | lo activation operand in place (no fp32 g, no split pass)
  __nv_bfloat16* gs = nullptr;
  int gnp = 0;
  if (v4 && g_tc.ws != nullptr && gemm_nn_tc_supported(a) && g_tc.bytes >= gemm_nn_tc_scratch_bytes(B, DIM, HID, N, false)) {
    gs = static_cast<__nv_bfloat16*>(gemm_nn_tc_act_operand(g_tc.ws, B, DIM, HID, N, false, &gnp));
    a.x_presplit = 1;
  }
  if (v4) gdfn_gate_fwd_v4_kernel<<<dim3(HID, B), 256, sm2, st>>>(s.tpre, params[P_FDW], g, gs, gnp, H, W);
  else gdfn_gate_fwd_kernel<<<dim3(HID, B), 256, sm2, st>>>(s.tpre, params[P_FDW], g, H, W);
  EMIP_CHECK_LAUNCH("gdfn_gate_fwd");
  a.res = s.y; a.res_stride_b = (long long)DIM * N; a.ldr = N;
  a.y = out; a.y_stride_b = (long long)DIM * N; a.ldy = N;
  return gemm_nn(a, st);
}

extern "C" int emip_injector_bwd(const float* x, const float* x1, const float* const* params, const void* saved,
                                 size_t saved_bytes_in, const float* dout, float* dx, float* dx1, float* const* dparams,
                                 void* workspace, size_t ws_bytes, int B, int H, int W, void* stream) {
  return emip_injector_bwd_ex(x, x1, params, saved, saved_bytes_in, dout, dx, dx1, dparams, workspace, ws_bytes, B, H, W, 0,
                              stream);
}

extern "C" int emip_injector_bwd_ex(const float* x, const float* x1, const float* const* params, const void* saved,
                                    size_t saved_bytes_in, const float* dout, float* dx, float* dx1, float* const* dparams,
                                    void* workspace, size_t ws_bytes, int B, int H, int W, int flags, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && x1 && params && saved && dout && dx && dx1 && dparams && workspace, "injector_bwd: null pointer");
  int rc = check_common("injector_bwd", B, H, W);
  if (rc) return rc;
  for (int i = 0; i < P_COUNT; ++i)
    EMIP_CHECK_ARG(params[i] != nullptr && dparams[i] != nullptr, "injector_bwd: parameter %d is NULL", i);
  cudaStream_t st = (cudaStream_t)stream;
  const int N = H * W;
  Saved s;
  if (saved_bytes_in < saved_bytes(B, N) || !carve_saved(const_cast<void*>(saved), saved_bytes_in, B, N, &s)) {
    emip_set_error("injector_bwd: saved buffer too small");
    return EMIP_ENOMEM;
  }
  if (ws_bytes < emip_injector_workspace(B, H, W)) {
    emip_set_error("injector_bwd: workspace too small");
    return EMIP_ENOMEM;
  }
  Carver w{static_cast<char*>(workspace), ws_bytes, true};
  // input- and weight-gradient GEMMs on the tensor cores unless the caller asks for exact fp32 (see emip_injector_fwd_ex)
  struct Scope {
    ~Scope() { g_tc = TcScope(); }
  } scope;
  if (!(flags & EMIP_FLAG_EXACT_FP32)) {
    g_tc.bytes = tc_scratch_bytes(B, N);
    g_tc.ws = reinterpret_cast<char*>(workspace) + (ws_bytes - g_tc.bytes) / 256 * 256;   // the tail of the workspace
  }
  float* g = w.take((size_t)B * HID * N);
  float* dg = w.take((size_t)B * HID * N);
  float* dtpre = w.take((size_t)B * HID2 * N);
  float* t128a = w.take((size_t)B * DIM * N);     // dn (gradient at a LayerNorm output)
  float* dy = w.take((size_t)B * DIM * N);
  float* dq = w.take((size_t)B * DIM * N);
  float* dk = w.take((size_t)B * DIM * N);
  float* dv = w.take((size_t)B * DIM * N);
  float* dqpre = w.take((size_t)B * DIM * N);
  float* dkvpre = w.take((size_t)B * 2 * DIM * N);
  float* wpart = w.take(nt_part_floats(B));
  float* Pm = w.take((size_t)B * DIM * DIM);
  float* dwo_part = w.take((size_t)B * DIM * DIM);
  float* dGs = w.take((size_t)B * HEADS * HD * HD);
  float* aq = w.take((size_t)B * DIM);
  float* ak = w.take((size_t)B * DIM);
  float* lg_part = w.take((size_t)B * DIM);
  float* lb_part = w.take((size_t)B * DIM);
  float* dwc_part = w.take((size_t)B * HID2 * 9);
  float* dtau_part = w.take((size_t)B * HEADS);
  if (!w.ok) { emip_set_error("injector_bwd: workspace carve failed"); return EMIP_ENOMEM; }
  const long long sN = (long long)DIM * N;

  // ---- GDFN: out = y + Wout g
  GemmNN a = {};
  a.B = B; a.M = HID; a.K = DIM; a.N = N; a.w = params[P_FOW]; a.ldw = HID; a.w_trans = 1;     // dg = Wout^T dout
  a.x = dout; a.x_stride_b = sN; a.ldx = N; a.y = dg; a.y_stride_b = (long long)HID * N; a.ldy = N;
  if ((rc = gemm_nn(a, st))) return rc;
  const size_t sm4 = plane_smem(H, W, 4);
  const bool v4 = (W & 3) == 0;
  if ((rc = v4 ? ensure_smem(gdfn_gate_bwd_v4_kernel, sm4) : ensure_smem(gdfn_gate_bwd_kernel, sm4))) return rc;
  if (v4) gdfn_gate_bwd_v4_kernel<<<dim3(HID, B), 256, sm4, st>>>(s.tpre, params[P_FDW], dg, dtpre, g, dwc_part, H, W);
  else gdfn_gate_bwd_kernel<<<dim3(HID, B), 256, sm4, st>>>(s.tpre, params[P_FDW], dg, dtpre, g, dwc_part, H, W);
  EMIP_CHECK_LAUNCH("gdfn_gate_bwd");
  if ((rc = reduce_batch(dwc_part, HID2 * 9, dparams[P_FDW], B, HID2 * 9, 0, st))) return rc;
  GemmNT t = {};
  t.B = B; t.M = DIM; t.K = HID; t.N = N;                                                        // dWout = dout g^T
  t.a = dout; t.a_stride_b = sN; t.lda = N; t.bm = g; t.b_stride_b = (long long)HID * N; t.ldb = N;
  t.c = wpart; t.c_stride_b = DIM * HID; t.ldc = HID;
  int ns = nt_split(B, DIM, HID);
  if ((rc = gemm_nt_split(t, ns, st))) return rc;
  if ((rc = reduce_batch(wpart, DIM * HID, dparams[P_FOW], B * ns, DIM * HID, 0, st))) return rc;
  t = {};
  t.B = B; t.M = HID2; t.K = DIM; t.N = N;                                                       // dWin = dtpre LN3(y)^T
  t.a = dtpre; t.a_stride_b = (long long)HID2 * N; t.lda = N; t.bm = s.y; t.b_stride_b = sN; t.ldb = N;
  t.mean = s.mean3; t.rstd = s.rstd3; t.gamma = params[P_N3W]; t.beta = params[P_N3B];
  t.c = wpart; t.c_stride_b = HID2 * DIM; t.ldc = DIM;
  ns = nt_split(B, HID2, DIM);
  if ((rc = gemm_nt_split(t, ns, st))) return rc;
  if ((rc = reduce_batch(wpart, HID2 * DIM, dparams[P_FIW], B * ns, HID2 * DIM, 0, st))) return rc;
  a = {};
  a.B = B; a.M = DIM; a.K = HID2; a.N = N; a.w = params[P_FIW]; a.ldw = DIM; a.w_trans = 1;      // dn3 = Win^T dtpre
  a.x = dtpre; a.x_stride_b = (long long)HID2 * N; a.ldx = N; a.y = t128a; a.y_stride_b = sN; a.ldy = N;
  if ((rc = gemm_nn(a, st))) return rc;
  ln_bwd_param_kernel<<<dim3(DIM, B), 256, 0, st>>>(t128a, s.y, s.mean3, s.rstd3, lg_part, lb_part, DIM, N);
  EMIP_CHECK_LAUNCH("ln_bwd_param 3");
  if ((rc = reduce_batch(lg_part, DIM, dparams[P_N3W], B, DIM, 0, st))) return rc;
  if ((rc = reduce_batch(lb_part, DIM, dparams[P_N3B], B, DIM, 0, st))) return rc;
  // dy = dout (residual) + LN3 backward
  ln_bwd_dx_kernel<<<dim3((N + LN_PX - 1) / LN_PX, B), LN_PX * LN_GR, 0, st>>>(t128a, s.y, s.mean3, s.rstd3, params[P_N3W], dout, dy, DIM, N);
  EMIP_CHECK_LAUNCH("ln_bwd_dx 3");

  // ---- MDTA: y = x + M[b] v,  M = Wo blockdiag(attn)
  t = {};
  t.B = B; t.M = DIM; t.K = DIM; t.N = N;                                                        // P = dy v^T
  t.a = dy; t.a_stride_b = sN; t.lda = N; t.bm = s.v; t.b_stride_b = sN; t.ldb = N;
  ns = nt_split(B, DIM, DIM);
  t.c = wpart; t.c_stride_b = DIM * DIM; t.ldc = DIM;
  if ((rc = gemm_nt_split(t, ns, st))) return rc;
  sum_splits_kernel<<<(B * DIM * DIM + 255) / 256, 256, 0, st>>>(wpart, Pm, B, ns, DIM * DIM);
  EMIP_CHECK_LAUNCH("sum_splits P");
  if ((rc = ensure_smem(mdta_attn_bwd_kernel, (size_t)2 * DIM * HD * sizeof(float)))) return rc;
  mdta_attn_bwd_kernel<<<dim3(B * HEADS, 5), 256, 2 * DIM * HD * sizeof(float), st>>>(Pm, s.attn, s.G, s.sq, s.sk, params[P_TEMP], params[P_POW], dGs, aq, ak,
                                                   dtau_part, dwo_part);
  EMIP_CHECK_LAUNCH("mdta_attn_bwd");
  if ((rc = reduce_batch(dwo_part, DIM * DIM, dparams[P_POW], B, DIM * DIM, 0, st))) return rc;
  if ((rc = reduce_batch(dtau_part, HEADS, dparams[P_TEMP], B, HEADS, 0, st))) return rc;
  a = {};
  a.B = B; a.M = DIM; a.K = DIM; a.N = N; a.w = s.M; a.w_stride_b = DIM * DIM; a.ldw = DIM; a.w_trans = 1;   // dv = M^T dy
  a.x = dy; a.x_stride_b = sN; a.ldx = N; a.y = dv; a.y_stride_b = sN; a.ldy = N;
  if ((rc = gemm_nn(a, st))) return rc;
  // dq = dG' k + a_q q ; dk = dG'^T q + a_k k      (per (sample, head): batch of 2B 64x64 problems)
  a = {};
  a.B = B * HEADS; a.M = HD; a.K = HD; a.N = N; a.w = dGs; a.w_stride_b = HD * HD; a.ldw = HD;
  a.x = s.k; a.x_stride_b = (long long)HD * N; a.ldx = N; a.y = dq; a.y_stride_b = (long long)HD * N; a.ldy = N;
  if ((rc = gemm_nn(a, st))) return rc;
  a.w_trans = 1; a.x = s.q; a.y = dk;
  if ((rc = gemm_nn(a, st))) return rc;
  row_axpy_kernel<<<dim3((N + 255) / 256, B * DIM), 256, 0, st>>>(aq, s.q, dq, N);
  row_axpy_kernel<<<dim3((N + 255) / 256, B * DIM), 256, 0, st>>>(ak, s.k, dk, N);
  EMIP_CHECK_LAUNCH("row_axpy");
  // depthwise backward
  const size_t sm2 = plane_smem(H, W, 2);
  if ((rc = v4 ? ensure_smem(dwconv_bwd_v4_kernel, sm2) : ensure_smem(dwconv_bwd_kernel, sm2))) return rc;
  if (v4) dwconv_bwd_v4_kernel<<<dim3(DIM, B), 256, sm2, st>>>(s.qpre, params[P_QDW], dq, nullptr, dqpre, dwc_part, DIM, DIM, H, W);
  else dwconv_bwd_kernel<<<dim3(DIM, B), 256, sm2, st>>>(s.qpre, params[P_QDW], dq, nullptr, dqpre, dwc_part, DIM, DIM, H, W);
  EMIP_CHECK_LAUNCH("dwconv_bwd q");
  if ((rc = reduce_batch(dwc_part, DIM * 9, dparams[P_QDW], B, DIM * 9, 0, st))) return rc;
  if (v4) dwconv_bwd_v4_kernel<<<dim3(2 * DIM, B), 256, sm2, st>>>(s.kvpre, params[P_KVDW], dk, dv, dkvpre, dwc_part, 2 * DIM, DIM, H, W);
  else dwconv_bwd_kernel<<<dim3(2 * DIM, B), 256, sm2, st>>>(s.kvpre, params[P_KVDW], dk, dv, dkvpre, dwc_part, 2 * DIM, DIM, H, W);
  EMIP_CHECK_LAUNCH("dwconv_bwd kv");
  if ((rc = reduce_batch(dwc_part, 2 * DIM * 9, dparams[P_KVDW], B, 2 * DIM * 9, 0, st))) return rc;
  // 1x1 convs + LayerNorms
  t = {};
  t.B = B; t.M = DIM; t.K = DIM; t.N = N;                                                        // dWq = dqpre LN1(x)^T
  t.a = dqpre; t.a_stride_b = sN; t.lda = N; t.bm = x; t.b_stride_b = sN; t.ldb = N;
  t.mean = s.mean1; t.rstd = s.rstd1; t.gamma = params[P_N1W]; t.beta = params[P_N1B];
  t.c = wpart; t.c_stride_b = DIM * DIM; t.ldc = DIM;
  ns = nt_split(B, DIM, DIM);
  if ((rc = gemm_nt_split(t, ns, st))) return rc;
  if ((rc = reduce_batch(wpart, DIM * DIM, dparams[P_QW], B * ns, DIM * DIM, 0, st))) return rc;
  t.M = 2 * DIM; t.a = dkvpre; t.a_stride_b = 2 * sN; t.bm = x1;                                  // dWkv = dkvpre LN2(x1)^T
  t.mean = s.mean2; t.rstd = s.rstd2; t.gamma = params[P_N2W]; t.beta = params[P_N2B];
  t.c_stride_b = 2 * DIM * DIM;
  ns = nt_split(B, 2 * DIM, DIM);
  if ((rc = gemm_nt_split(t, ns, st))) return rc;
  if ((rc = reduce_batch(wpart, 2 * DIM * DIM, dparams[P_KVW], B * ns, 2 * DIM * DIM, 0, st))) return rc;
  a = {};
  a.B = B; a.M = DIM; a.K = DIM; a.N = N; a.w = params[P_QW]; a.ldw = DIM; a.w_trans = 1;        // dn1 = Wq^T dqpre
  a.x = dqpre; a.x_stride_b = sN; a.ldx = N; a.y = t128a; a.y_stride_b = sN; a.ldy = N;
  if ((rc = gemm_nn(a, st))) return rc;
  ln_bwd_param_kernel<<<dim3(DIM, B), 256, 0, st>>>(t128a, x, s.mean1, s.rstd1, lg_part, lb_part, DIM, N);
  EMIP_CHECK_LAUNCH("ln_bwd_param 1");
  if ((rc = reduce_batch(lg_part, DIM, dparams[P_N1W], B, DIM, 0, st))) return rc;
  if ((rc = reduce_batch(lb_part, DIM, dparams[P_N1B], B, DIM, 0, st))) return rc;
  // dx = dy (residual x + ...) + LN1 backward
  ln_bwd_dx_kernel<<<dim3((N + LN_PX - 1) / LN_PX, B), LN_PX * LN_GR, 0, st>>>(t128a, x, s.mean1, s.rstd1, params[P_N1W], dy, dx, DIM, N);
  EMIP_CHECK_LAUNCH("ln_bwd_dx 1");
  a.K = 2 * DIM; a.w = params[P_KVW]; a.x = dkvpre; a.x_stride_b = 2 * sN;                         // dn2 = Wkv^T dkvpre
  if ((rc = gemm_nn(a, st))) return rc;
  ln_bwd_param_kernel<<<dim3(DIM, B), 256, 0, st>>>(t128a, x1, s.mean2, s.rstd2, lg_part, lb_part, DIM, N);
  EMIP_CHECK_LAUNCH("ln_bwd_param 2");
  if ((rc = reduce_batch(lg_part, DIM, dparams[P_N2W], B, DIM, 0, st))) return rc;
  if ((rc = reduce_batch(lb_part, DIM, dparams[P_N2B], B, DIM, 0, st))) return rc;
  ln_bwd_dx_kernel<<<dim3((N + LN_PX - 1) / LN_PX, B), LN_PX * LN_GR, 0, st>>>(t128a, x1, s.mean2, s.rstd2, params[P_N2W], nullptr, dx1, DIM, N);
  EMIP_CHECK_LAUNCH("ln_bwd_dx 2");
  return EMIP_OK;
}
