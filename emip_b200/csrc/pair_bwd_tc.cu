// Backward of the pairwise-softmax family (a1 global matching, a2 flow-propagation attention) on the 5th-gen
// tensor cores.  Same mathematics as pair_bwd_simt_kernel (pair_simt.cu), which stays the exact-fp32 reference path:
//   S = X Y^T / sqrt(C)            (recomputed tile by tile, never stored)
//   W[r,c] = e^{S - L1[r]} (u[r].t[c] - u0[r]) + e^{S - L2[c]} (w[c].t2[r] - w0[c]) + E[r,c]
//   dX[r,:] = sum_c W[r,c] Y[c,:] / sqrt(C)
// i.e. what autograd derives for matching.py:16-39 (both softmax directions and the dcorr term in one pass) and
// for transformer.py:528-531 with a detached value.
//
// One CTA = one 128-row tile of X, persistent over (problem, row tile) items, 576 threads:
//   warps 0..15 gradient-tile math + epilogue (thread <-> TMEM lane <-> row; warp w: lane quarter w%4, 32-column
//               part w/4 of every key tile).  The tile math is ~25 instructions per element with two exponentials,
//               far more than the forward's softmax, so it -- not the UMMAs -- bounds this kernel: 16 warps.
//   warp 16     TMA producer: X tile (token-major hi|lo) once per item, then per 128-column key tile four
//               token-major chunks of Y (for S) and four channel-major chunks of Y (for dX) through one 8-stage ring
//   warp 17     UMMA issuer
// TMEM (512 columns): S double-buffered [0,256) | W as bf16 hi [256,320) + lo [320,384) | dX accumulator [384,512).
//   UMMA-1 (SS): S  = X.hi Y.hi^T + X.lo Y.hi^T + X.hi Y.lo^T                         (as in the forward kernel)
//   UMMA-2 (TS): dX += W.hi Y.hi + W.lo Y.hi + W.hi Y.lo   with A = W read from TMEM, B = channel-major Y chunks
// W is split into bf16 hi + lo like the features, so every product keeps ~16 mantissa bits (gradients agree with the
// fp32 path to ~1e-5).  The issuer runs UMMA-1 of tile t+1 ahead of UMMA-2 of tile t, so the tensor pipe recomputes
// the next S tile while the math warps turn the current one into W.
#include "common.cuh"
#include "pair_common.cuh"
#include "tc_common.cuh"
#include "pair_bwd_tc.cuh"

namespace {
using namespace tc;

constexpr int TN = 128;                 // key columns per tile
constexpr int CH_ELEMS = 64;
constexpr int CHUNK_BYTES = TM * 128;   // 16 KB
constexpr int STAGES = 8;
constexpr int NMATH = 16;                // math warps
constexpr int NTHREADS = (NMATH + 2) * 32;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_W = 256, COL_DX = 384;

constexpr int OFF_X = 0;                                  // 4 chunks: hi[0:64] hi[64:128] lo[0:64] lo[64:128]
constexpr int OFF_RING = OFF_X + 4 * CHUNK_BYTES;
constexpr int OFF_TAB = OFF_RING + STAGES * CHUNK_BYTES;  // [2 buffers][6][TN] floats: tx ty L2' wx wy w0
constexpr int OFF_BAR = OFF_TAB + 2 * 6 * TN * 4;
constexpr int NBAR = 2 + 2 * STAGES + 4 + 4;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

struct BParams {
  PairBwdTcArgs a;
  float inv_sqrt_c;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);      // .x = first (lower address) element
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(NTHREADS, 1)
pair_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y1,
                   const __grid_constant__ CUtensorMap map_y2, BParams bp) {
  const PairBwdTcArgs& p = bp.a;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + OFF_BAR;
  const uint32_t x_full = bar0, x_empty = bar0 + 8;
  auto r_full = [&](int s) { return bar0 + 16 + 8 * s; };
  auto r_empty = [&](int s) { return bar0 + 16 + 8 * (STAGES + s); };
  auto s_full = [&](int b) { return bar0 + 16 + 8 * (2 * STAGES + b); };
  auto s_empty = [&](int b) { return bar0 + 16 + 8 * (2 * STAGES + 2 + b); };
  const uint32_t w_full = bar0 + 16 + 8 * (2 * STAGES + 4), w_empty = w_full + 8, dx_full = w_full + 16, dx_empty = w_full + 24;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrt = (p.nr + TM - 1) / TM;
  const int nkt = (p.nc + TN - 1) / TN;
  // an item is (problem, row tile, key split): split ks covers key tiles [ks*nkt/ns, (ks+1)*nkt/ns) and writes its own
  // partial dX (stacked by split index; summed by the caller) -- used when nb * nrt alone cannot fill the SMs
  const int ns = p.ksplit > 1 ? p.ksplit : 1;
  const int n_items = p.nb * nrt * ns;

  if (threadIdx.x == 0) {
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(r_full(s), 1); mbar_init(r_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(s_empty(b), NMATH * 32); }
    mbar_init(w_full, NMATH * 32);
    mbar_init(w_empty, 1);
    mbar_init(dx_full, 1);
    mbar_init(dx_empty, NMATH * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NMATH + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + OFF_TMEM),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == NMATH) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0, it = 0;
    auto push = [&](const CUtensorMap* map, int c0, int c1, int c2) {
      mbar_wait(r_empty(stage), phase ^ 1);
      if (leader) {
        mbar_expect_tx(r_full(stage), CHUNK_BYTES);
        tma_load_3d(sbase + OFF_RING + stage * CHUNK_BYTES, map, r_full(stage), c0, c1, c2);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int ks = item % ns, rt = (item / ns) % nrt, prob = item / (ns * nrt);
      const int kb = ks * nkt / ns, ke = (ks + 1) * nkt / ns;
      const int xb = p.x_base + prob, yb = p.y_base + prob;
      mbar_wait(x_empty, (it & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(x_full, 4 * CHUNK_BYTES);
        for (int c = 0; c < 4; ++c) tma_load_3d(sbase + OFF_X + c * CHUNK_BYTES, &map_x, x_full, c * CH_ELEMS, rt * TM, xb);
      }
      // ring order = consumption order of the issuer: Y1(0), then per tile { Y1(t+1), Y2(t) }
      for (int c = 0; c < 4; ++c) push(&map_y1, c * CH_ELEMS, kb * TN, yb);
      for (int t = kb; t < ke; ++t) {
        if (t + 1 < ke)
          for (int c = 0; c < 4; ++c) push(&map_y1, c * CH_ELEMS, (t + 1) * TN, yb);
        for (int c = 0; c < 4; ++c)       // hi keys[0:64], hi keys[64:128], lo keys[0:64], lo keys[64:128]
          push(&map_y2, t * TN + (c & 1) * CH_ELEMS, (c >> 1) * 128, yb);
      }
    }
    __syncwarp();
  } else if (warp == NMATH + 1) {
    // ===================== UMMA issuer =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0, it = 0, tile = 0;
    const int n_tail = ((p.nc - (nkt - 1) * TN) + 15) & ~15;
    const uint32_t idesc_full = make_idesc(TN), idesc_tail = make_idesc(n_tail);
    const uint64_t xd = make_kmajor_sw128_desc(sbase + OFF_X);
    const uint32_t t_w = tmem_base + COL_W, t_dx = tmem_base + COL_DX;
    auto mma1 = [&](uint32_t T, bool tail) {              // S tile T -> TMEM buffer T & 1
      const int buf = T & 1;
      mbar_wait(s_empty(buf), ((T >> 1) & 1) ^ 1);
      const uint32_t idesc = tail ? idesc_tail : idesc_full;
      const uint32_t d = tmem_base + (uint32_t)(buf * TN);
      for (int c = 0; c < 4; ++c) {
        mbar_wait(r_full(stage), phase);
        tc_fence_after();
        if (leader) {
          const uint64_t kd = make_kmajor_sw128_desc(sbase + OFF_RING + stage * CHUNK_BYTES);
          const uint64_t a_hi = xd + (uint64_t)((c & 1) * (CHUNK_BYTES >> 4));
          const uint64_t a_lo = a_hi + (uint64_t)(2 * (CHUNK_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, a_hi + 2 * k, kd + 2 * k, idesc, (c | k) ? 1u : 0u);
          if (c < 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, a_lo + 2 * k, kd + 2 * k, idesc, 1u);
          }
          umma_commit(r_empty(stage));
          if (c == 3) umma_commit(s_full(buf));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    auto mma2 = [&](uint32_t T, bool first, bool last) {  // dX += W(T) Y2(T)
      mbar_wait(w_full, T & 1);
      if (first) mbar_wait(dx_empty, (it & 1) ^ 1);       // the epilogue of the previous item has drained dX
      for (int c = 0; c < 4; ++c) {
        mbar_wait(r_full(stage), phase);
        tc_fence_after();
        if (leader) {
          const uint64_t bd = make_kmajor_sw128_desc(sbase + OFF_RING + stage * CHUNK_BYTES);
          const int j = c & 1;                            // which 64 keys of the tile
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ks = (uint32_t)(8 * (4 * j + k));
            umma_bf16_ts(t_dx, t_w + ks, bd + 2 * k, idesc_full, (first && c == 0 && k == 0) ? 0u : 1u);   // W.hi Y.(hi|lo)
            if (c < 2) umma_bf16_ts(t_dx, t_w + 64 + ks, bd + 2 * k, idesc_full, 1u);                      // W.lo Y.hi
          }
          umma_commit(r_empty(stage));
          if (c == 3) {
            umma_commit(w_empty);
            if (last) umma_commit(dx_full);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int ks = item % ns;
      const int kb = ks * nkt / ns, ke = (ks + 1) * nkt / ns;
      mbar_wait(x_full, it & 1);
      tc_fence_after();
      mma1(tile, kb == nkt - 1);
      for (int t = kb; t < ke; ++t) {
        if (t + 1 < ke) mma1(tile + (t - kb) + 1, t + 1 == nkt - 1);
        mma2(tile + (t - kb), t == kb, t == ke - 1);
      }
      tile += ke - kb;
      if (leader) umma_commit(x_empty);
      __syncwarp();
    }
    __syncwarp();
  } else {
    // ===================== gradient-tile math + epilogue =====================
    const int quarter = warp & 3, part = warp >> 2;        // TMEM lane quarter; 32-column part of every key tile
    const int st = threadIdx.x;                            // 0..511
    float* tab = reinterpret_cast<float*>(smem + OFF_TAB);
    const uint32_t tab_u32 = sbase + OFF_TAB;
    const float LOG2E = 1.4426950408889634f;
    const float c2 = LOG2E * bp.inv_sqrt_c;
    const bool term1 = p.l1 != nullptr, term2 = p.l2 != nullptr, extra = p.e != nullptr;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int cb = part * 32;                              // first column (within the tile) of my chunk
    // E rows that are contiguous along the columns can be read with 128-bit loads
    const bool e_vec = extra && p.e_stride_c == 1 && (p.e_stride_r % 4) == 0 && (p.e_stride_b % 4) == 0 &&
                       (reinterpret_cast<uintptr_t>(p.e) % 16) == 0;
    uint32_t tile = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int ks = item % ns, rt = (item / ns) % nrt, prob = item / (ns * nrt);
      const int kb = ks * nkt / ns, ke = (ks + 1) * nkt / ns;
      const int row = rt * TM + quarter * 32 + lane;
      const bool row_ok = row < p.nr;
      float rL1 = 0.f, rux = 0.f, ruy = 0.f, ru0 = 0.f, rtx = 0.f, rty = 0.f;
      if (row_ok && term1) {
        rL1 = __ldg(p.l1 + (size_t)prob * p.nr + row) * LOG2E;
        rux = __ldg(p.u + ((size_t)prob * 2 + 0) * p.nr + row);
        ruy = __ldg(p.u + ((size_t)prob * 2 + 1) * p.nr + row);
        ru0 = __ldg(p.u0 + (size_t)prob * p.nr + row);
      }
      if (row_ok && term2) {
        rtx = __ldg(p.t2 + (size_t)prob * p.t2_stride_b + row);
        rty = __ldg(p.t2 + (size_t)prob * p.t2_stride_b + p.nr + row);
      }
      const float* E = extra ? p.e + (size_t)prob * p.e_stride_b + (size_t)row * p.e_stride_r : nullptr;
      for (int kt = kb; kt < ke; ++kt, ++tile) {
        const int buf = tile & 1;
        const int col_base = kt * TN;
        // per-column terms of this key tile -> smem (double-buffered; one barrier per tile orders writes and reads)
        float* tb = tab + buf * 6 * TN;
        for (int i = st; i < 6 * TN; i += NMATH * 32) {
          const int k = i / TN, c = i % TN, col = col_base + c;
          float v = 0.f;
          if (col < p.nc) {
            if (k < 2) { if (term1) v = __ldg(p.t + (size_t)prob * p.t_stride_b + (size_t)k * p.nc + col); }
            else if (term2) {
              if (k == 2) v = __ldg(p.l2 + (size_t)prob * p.nc + col) * LOG2E;
              else if (k == 3) v = __ldg(p.w + ((size_t)prob * 2 + 0) * p.nc + col);
              else if (k == 4) v = __ldg(p.w + ((size_t)prob * 2 + 1) * p.nc + col);
              else v = __ldg(p.w0 + (size_t)prob * p.nc + col);
            }
          }
          tb[i] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NMATH * 32) : "memory");
        mbar_wait(s_full(buf), (tile >> 1) & 1);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32_async(tmem_base + lane_base + (uint32_t)(buf * TN + cb), r);
        tmem_wait(r);
        tc_fence_before();
        mbar_arrive(s_empty(buf));                          // my part of the S tile is in registers
        const uint32_t tbu = tab_u32 + (uint32_t)((buf * 6 * TN + cb) * 4);
        uint32_t whi[16], wlo[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float w4[4] = {0.f, 0.f, 0.f, 0.f};
          if (term1) {
            const float4 tx = lds128(tbu + 16 * q), ty = lds128(tbu + 4 * TN + 16 * q);
            const float txa[4] = {tx.x, tx.y, tx.z, tx.w}, tya[4] = {ty.x, ty.y, ty.z, ty.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              w4[e] = ex2f(fmaf(__uint_as_float(r[4 * q + e]), c2, -rL1)) * (fmaf(rux, txa[e], ruy * tya[e]) - ru0);
          }
          if (term2) {
            const float4 l2 = lds128(tbu + 8 * TN + 16 * q), wx = lds128(tbu + 12 * TN + 16 * q),
                         wy = lds128(tbu + 16 * TN + 16 * q), w0 = lds128(tbu + 20 * TN + 16 * q);
            const float l2a[4] = {l2.x, l2.y, l2.z, l2.w}, wxa[4] = {wx.x, wx.y, wx.z, wx.w},
                        wya[4] = {wy.x, wy.y, wy.z, wy.w}, w0a[4] = {w0.x, w0.y, w0.z, w0.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              w4[e] = fmaf(ex2f(fmaf(__uint_as_float(r[4 * q + e]), c2, -l2a[e])), fmaf(wxa[e], rtx, wya[e] * rty) - w0a[e], w4[e]);
          }
          if (extra) {                                      // direct term (dcorr): 16 resident warps hide the load latency
            const int colq = col_base + cb + 4 * q;
            if (row_ok && e_vec && colq + 4 <= p.nc) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(E + colq));
              w4[0] += v.x; w4[1] += v.y; w4[2] += v.z; w4[3] += v.w;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (row_ok && colq + e < p.nc) w4[e] += __ldg(E + (size_t)(colq + e) * p.e_stride_c);
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (!(row_ok && col_base + cb + 4 * q + e < p.nc)) w4[e] = 0.f;
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const float a0 = w4[2 * e2], a1 = w4[2 * e2 + 1];
            const float h0 = __bfloat162float(__float2bfloat16_rn(a0)), h1 = __bfloat162float(__float2bfloat16_rn(a1));
            whi[2 * q + e2] = pack_bf16x2(h0, h1);
            wlo[2 * q + e2] = pack_bf16x2(a0 - h0, a1 - h1);
          }
        }
        // W buffer is free once UMMA-2 of the previous tile has retired
        mbar_wait(w_empty, (tile & 1) ^ 1);
        tc_fence_after();
        const uint32_t t_whi = tmem_base + lane_base + COL_W + (uint32_t)(part * 16);
        tmem_st16(t_whi, whi);
        tmem_st16(t_whi + 64, wlo);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(w_full);
      }
      // ---- epilogue: dX tile -> global (scaled by 1/sqrt(C): the chain through S = X Y^T / sqrt(C))
      mbar_wait(dx_full, it & 1);
      tc_fence_after();
      float* DX = p.dx + ((size_t)ks * p.nb + prob) * p.nr * 128;
      {
        uint32_t r[32];
        tmem_ld32_async(tmem_base + lane_base + COL_DX + (uint32_t)cb, r);
        tmem_wait(r);
        if (row_ok) {
          if (p.dx_layout == EMIP_LAYOUT_NC) {
            float4* dst = reinterpret_cast<float4*>(DX + (size_t)row * 128 + cb);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              dst[q] = make_float4(__uint_as_float(r[4 * q]) * bp.inv_sqrt_c, __uint_as_float(r[4 * q + 1]) * bp.inv_sqrt_c,
                                   __uint_as_float(r[4 * q + 2]) * bp.inv_sqrt_c, __uint_as_float(r[4 * q + 3]) * bp.inv_sqrt_c);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) DX[(size_t)(cb + i) * p.nr + row] = __uint_as_float(r[i]) * bp.inv_sqrt_c;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(dx_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NMATH + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---- channel-major operand split: fp32 -> bf16 hi rows | bf16 lo rows --------------------------------
// dst [nb][256][ld]: rows 0..127 = hi(channel), 128..255 = lo(channel), keys contiguous.
// layout CN: src [nb][128][n] is already channel-major (element-wise); layout NC: src [nb][n][128] is transposed.
__global__ void __launch_bounds__(256)
split_chn_kernel(const float* __restrict__ src, const float* __restrict__ src2, __nv_bfloat16* __restrict__ dst, int nb, int n,
                 long long ld, int layout) {
  const int b = blockIdx.z;
  const float* s = b < nb ? src + (size_t)b * 128 * n : src2 + (size_t)(b - nb) * 128 * n;
  __nv_bfloat16* d = dst + (size_t)b * 256 * ld;
  if (layout == EMIP_LAYOUT_CN) {
    const int c = blockIdx.y;
    if ((n & 7) == 0 && (reinterpret_cast<uintptr_t>(s) & 15) == 0) {
      // eight consecutive keys per thread: two 128-bit loads, one 128-bit store each for the hi and the lo row
      // (the scalar loop below wrote 2-byte elements: 61 us for the 63 MB of an a2 backward)
      for (int k = (blockIdx.x * blockDim.x + threadIdx.x) * 8; k < n; k += gridDim.x * blockDim.x * 8) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(s + (size_t)c * n + k));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(s + (size_t)c * n + k + 4));
        const float v[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
          const __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
          const __nv_bfloat162 ll = __halves2bfloat162(__float2bfloat16_rn(v[2 * i] - __bfloat162float(h0)),
                                                        __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1)));
          h[i] = *reinterpret_cast<const uint32_t*>(&hh);
          l[i] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        *reinterpret_cast<uint4*>(d + (size_t)c * ld + k) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(d + (size_t)(128 + c) * ld + k) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      return;
    }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
      const float v = __ldg(s + (size_t)c * n + k);
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      d[(size_t)c * ld + k] = h;
      d[(size_t)(128 + c) * ld + k] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  } else {
    // 32 keys x 128 channels per block through smem
    __shared__ float t[32][129];
    const int k0 = (blockIdx.y * gridDim.x + blockIdx.x) * 32;
    if (k0 >= n) return;
    for (int i = threadIdx.x; i < 32 * 128; i += 256) {
      const int k = i / 128, c = i % 128;
      t[k][c] = (k0 + k < n) ? __ldg(s + (size_t)(k0 + k) * 128 + c) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 128; i += 256) {
      const int c = i / 32, k = i % 32;
      if (k0 + k < n) {
        const float v = t[k][c];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        d[(size_t)c * ld + k0 + k] = h;
        d[(size_t)(128 + c) * ld + k0 + k] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
  }
}

}  // namespace

bool pair_bwd_tc_supported(int nr, int nc, int c) { return c == 128 && nr >= 1 && nc >= 16; }

long long pair_bwd_tc_chn_ld(int n) { return ((long long)n + 7) / 8 * 8; }
size_t pair_bwd_tc_chn_bytes(int nb, int n) { return emip_align_up((size_t)nb * 256 * pair_bwd_tc_chn_ld(n) * 2, 1024); }

int pair_bwd_tc_split_chn(const float* src, const float* src2, void* dst, int nb, int n, int layout, cudaStream_t st) {
  if (nb == 0 || n == 0) return EMIP_OK;
  const int tot = src2 ? 2 * nb : nb;
  if (layout == EMIP_LAYOUT_CN) {
    dim3 grid((n + 255) / 256, 128, tot);
    split_chn_kernel<<<grid, 256, 0, st>>>(src, src2, static_cast<__nv_bfloat16*>(dst), nb, n, pair_bwd_tc_chn_ld(n), layout);
  } else {
    const int blocks = (n + 31) / 32;
    dim3 grid(blocks, 1, tot);
    split_chn_kernel<<<grid, 256, 0, st>>>(src, src2, static_cast<__nv_bfloat16*>(dst), nb, n, pair_bwd_tc_chn_ld(n), layout);
  }
  EMIP_CHECK_LAUNCH("pair_bwd_tc_split_chn");
  return EMIP_OK;
}

int pair_bwd_tc(const PairBwdTcArgs& a, cudaStream_t st) {
  if (a.nb == 0 || a.nr == 0) return EMIP_OK;
  if (!pair_bwd_tc_supported(a.nr, a.nc, 128)) { emip_set_error("pair_bwd_tc: unsupported shape"); return EMIP_ENOSYS; }
  CUtensorMap mx, my1, my2;
  int rc;
  if ((rc = make_bf16_map(&mx, a.tok_split, 256, (uint64_t)a.nr, (uint64_t)a.n_split, 512, (uint64_t)a.nr * 512))) return rc;
  if ((rc = make_bf16_map(&my1, a.tok_split_y, 256, (uint64_t)a.nc, (uint64_t)a.n_split, 512, (uint64_t)a.nc * 512))) return rc;
  const uint64_t ld = (uint64_t)pair_bwd_tc_chn_ld(a.nc);
  if ((rc = make_bf16_map(&my2, a.chn_split_y, (uint64_t)a.nc, 256, (uint64_t)a.n_split, ld * 2, ld * 2 * 256))) return rc;
  if (int rc__ = emip_func_max_smem((const void*)(pair_bwd_tc_kernel), SMEM_BYTES)) return rc__;
  BParams bp;
  bp.a = a;
  bp.inv_sqrt_c = 1.0f / a.sqrt_c;
  const int nrt = (a.nr + TM - 1) / TM;
  int grid = a.nb * nrt * (a.ksplit > 1 ? a.ksplit : 1);
  if (grid > emip_num_sms()) grid = emip_num_sms();
  pair_bwd_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(mx, my1, my2, bp);
  EMIP_CHECK_LAUNCH("pair_bwd_tc");
  return EMIP_OK;
}
