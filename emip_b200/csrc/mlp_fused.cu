// The feed-forward network of the FeatureTransformer's cross-attention layer as ONE kernel:
//   y = res + LayerNorm(GELU([source | message] W1^T) W2^T) * gamma + beta        (transformer.py:175-176, :180)
// with the 1024-wide hidden rows never leaving the SM.  As two gemm_tc launches (r3) the hidden tensor went out to HBM as a
// bf16 hi | lo operand and came back in: 2 x 1.02 GB per call at 64 pairs, a third of the FeatureTransformer's HBM traffic,
// and mlp[2] ran at the HBM rate of that read-back (r4h: 362 + 221 us per block, unthrottled).
//
// CTA = one 128-row tile of token rows, persistent over the row tiles, 576 threads (the role layout of attn_tc.cu):
//   warps 0..15 GELU + operand split, then the LayerNorm epilogue (thread <-> TMEM lane <-> row; warp w: lane quarter w % 4,
//               32-column part w / 4 of every 128-column block)
//   warp 16     TMA producer: the row tile X (4 K chunks x hi, lo = 128 KB, resident for the 8 hidden blocks), then per
//               hidden block j the W1 rows of the block (4 K chunks x (hi, lo)) and the W2 columns of the block (2 K chunks
//               x (hi, lo)) through a 6-stage ring of 16 KB tiles, in the issuer's consumption order
//   warp 17     UMMA issuer
// TMEM (512 columns): S double-buffered [0, 256) | H as bf16 hi [256, 320) + lo [320, 384) | OUT accumulator [384, 512).
//   UMMA-1 (SS): S_j  = X.hi W1_j.hi^T + X.lo W1_j.hi^T + X.hi W1_j.lo^T         (K = 256, the 3-term bf16 split)
//   math warps : H_j  = GELU(S_j) -> bf16 hi | lo, written back to TMEM (tcgen05.st) as the A operand of UMMA-2
//   UMMA-2 (TS): OUT += H_j.hi W2_j.hi^T + H_j.lo W2_j.hi^T + H_j.hi W2_j.lo^T   (A read from TMEM: no shared-memory copy of H)
// The issuer runs UMMA-1 of block j + 1 ahead of UMMA-2 of block j, so the GELU of block j overlaps the first GEMM of the next.
// S_j accumulates in the order of gemm_tc_kernel (per 64-wide K chunk: hi.hi, lo.hi, hi.lo) and H is split with the same
// round-to-nearest conversions as its epilogue, so the hidden operand is bit-identical to the two-launch path; OUT sums the same
// products in a different order (per 64-wide K half: hi.hi / lo.hi interleaved per K16 step, then hi.lo), ~1e-7 relative.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/emip_b200.h"
#include "mlp_fused.cuh"
#include "attn_tc.cuh"

namespace {
using namespace tc;

constexpr int TN = 128;                  // hidden columns per block = UMMA N of both GEMMs
constexpr int CHUNK_BYTES = TM * 128;    // 16 KB: [128 rows][64 bf16], 128-byte swizzled
constexpr int KX = 256, NKC = KX / 64;   // contraction width of the first layer, in 64-element chunks
constexpr int STAGES = 6;                // ring of 16 KB tiles; always advanced in pairs (hi, lo) / (k 0..63, k 64..127)
constexpr int NMATH = 16;
constexpr int NTHREADS = (NMATH + 2) * 32;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_H = 256, COL_O = 384;

constexpr int OFF_X = 0;                                      // 8 tiles: hi chunk 0..3, lo chunk 0..3
constexpr int OFF_RING = OFF_X + 2 * NKC * CHUNK_BYTES;
constexpr int OFF_BAR = OFF_RING + STAGES * CHUNK_BYTES;
constexpr int NBAR = 2 + 2 * STAGES + 4 + 4 + 2;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
static_assert(STAGES % 2 == 0, "ring stages are consumed in adjacent pairs");

struct MParams {
  MlpFusedArgs a;
  int n_tiles, nj;
  unsigned long long* prof;   // diagnostics (tools/mlp_roles.py; the buffer of emip_attn_tc_set_profile_buffer): [grid][16] wait / work cycles
};

template <bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                 const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2,
                 const __grid_constant__ MParams mp) {
  const MlpFusedArgs& p = mp.a;
  pdl_trigger();                                          // common.cuh: the next kernel's prologue may overlap this kernel's tail
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + OFF_BAR;
  const uint32_t x_full = bar0, x_empty = bar0 + 8;
  auto r_full = [&](int s) { return bar0 + 16 + 8 * s; };
  auto r_empty = [&](int s) { return bar0 + 16 + 8 * (STAGES + s); };
  auto s_full = [&](int b) { return bar0 + 16 + 8 * (2 * STAGES + b); };
  auto s_empty = [&](int b) { return bar0 + 16 + 8 * (2 * STAGES + 2 + b); };
  const uint32_t h_full = bar0 + 16 + 8 * (2 * STAGES + 4), h_empty = h_full + 8, o_full = h_full + 16, o_empty = h_full + 24;
  // x_full = the hi half of X; the lo half has its own barrier because its shared memory doubles as the staging tiles of the
  // LayerNorm epilogue (16 warps x 4 KB: there is no other room) and is reloaded only when those are done (stg_free)
  const uint32_t xl_full = h_full + 32, stg_free = h_full + 40;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nj = mp.nj, n_items = mp.n_tiles;

  if (threadIdx.x == 0) {
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(r_full(s), 1); mbar_init(r_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(s_empty(b), NMATH * 32); }
    mbar_init(h_full, NMATH * 32);
    mbar_init(h_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, NMATH * 32);
    mbar_init(xl_full, 1);
    mbar_init(stg_free, NMATH * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NMATH + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + OFF_TMEM), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                             // everything below reads what the kernel in front of this one wrote

  if (warp == NMATH) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0, it = 0;
    long long w_re = 0, w_xe = 0;
    const long long t_begin = PROF ? clock64() : 0;
    bool xlo_pending = false;
    int xlo_item = 0;
    uint32_t xlo_parity = 0;
    auto maybe_load_xlo = [&]() {        // never blocks: the ring must keep moving while the LayerNorm epilogue holds the lo half
      if (xlo_pending && mbar_try_wait(stg_free, xlo_parity)) {
        if (leader) {
          mbar_expect_tx(xl_full, NKC * CHUNK_BYTES);
          for (int kc = 0; kc < NKC; ++kc)
            tma_load_3d(sbase + OFF_X + (NKC + kc) * CHUNK_BYTES, &map_x_lo, xl_full, kc * 64, xlo_item * TM, 0);
        }
        xlo_pending = false;
      }
    };
    auto push = [&](const CUtensorMap* map, int c0, int c1) {
      if (xlo_pending) {
        const long long t0 = PROF ? clock64() : 0;
        while (!mbar_try_wait(r_empty(stage), phase ^ 1)) maybe_load_xlo();
        maybe_load_xlo();
        if (PROF) w_re += clock64() - t0;
      }
      prof_add<PROF>(w_re, mbar_wait(r_empty(stage), phase ^ 1));
      if (leader) {
        mbar_expect_tx(r_full(stage), CHUNK_BYTES);
        tma_load_3d(sbase + OFF_RING + stage * CHUNK_BYTES, map, r_full(stage), c0, c1, 0);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    };
    auto push_w1 = [&](int j) {          // rows j*128.. of W1: per K chunk the hi tile, then the lo tile
      for (int kc = 0; kc < NKC; ++kc) {
        push(&map_w1, kc * 64, j * TN);
        push(&map_w1, KX + kc * 64, j * TN);
      }
    };
    auto push_w2 = [&](int j) {          // columns j*128.. of W2: per 64-wide K half the hi tile, then the lo tile (like W1)
      push(&map_w2, j * TN, 0);
      push(&map_w2, p.hid + j * TN, 0);
      push(&map_w2, j * TN + 64, 0);
      push(&map_w2, p.hid + j * TN + 64, 0);
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      prof_add<PROF>(w_xe, mbar_wait(x_empty, (it & 1) ^ 1));
      if (leader) {
        mbar_expect_tx(x_full, NKC * CHUNK_BYTES);
        for (int kc = 0; kc < NKC; ++kc) tma_load_3d(sbase + OFF_X + kc * CHUNK_BYTES, &map_x_hi, x_full, kc * 64, item * TM, 0);
      }
      // the lo half once the LayerNorm epilogue of the previous row tile has read its staging tiles back (first tile: at once)
      xlo_pending = true; xlo_item = item; xlo_parity = (it & 1) ^ 1;
      maybe_load_xlo();
      // ring order = consumption order of the issuer: W1(0), then per block { W1(j + 1), W2(j) }
      push_w1(0);
      for (int j = 0; j < nj; ++j) {
        if (j + 1 < nj) push_w1(j + 1);
        push_w2(j);
      }
      while (xlo_pending) maybe_load_xlo();          // (cannot happen: the issuer needs the lo half long before the ring drains)
    }
    if (PROF && mp.prof && leader) {
      unsigned long long* o = mp.prof + blockIdx.x * 16;
      o[13] = w_re + w_xe; o[15] = clock64() - t_begin;
    }
    __syncwarp();
  } else if (warp == NMATH + 1) {
    // ===================== UMMA issuer =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0, it = 0, tile = 0;
    const uint32_t idesc = make_idesc(TN);
    const uint32_t t_h = tmem_base + COL_H, t_o = tmem_base + COL_O;
    long long w_se = 0, w_rf = 0, w_rf2 = 0, w_hf = 0, w_oe = 0, w_xf = 0;
    const long long t_begin = PROF ? clock64() : 0;
    auto mma1 = [&](uint32_t T, bool first_of_tile) {      // S block T -> TMEM buffer T & 1
      const int buf = T & 1;
      prof_add<PROF>(w_se, mbar_wait(s_empty(buf), ((T >> 1) & 1) ^ 1));
      const uint32_t d = tmem_base + (uint32_t)(buf * TN);
      for (int kc = 0; kc < NKC; ++kc) {
        prof_add<PROF>(w_rf, mbar_wait(r_full(stage), phase));
        prof_add<PROF>(w_rf, mbar_wait(r_full(stage + 1), phase));
        tc_fence_after();
        const uint64_t xh = make_kmajor_sw128_desc(sbase + OFF_X + kc * CHUNK_BYTES);
        const uint64_t xl = make_kmajor_sw128_desc(sbase + OFF_X + (NKC + kc) * CHUNK_BYTES);
        const uint64_t wh = make_kmajor_sw128_desc(sbase + OFF_RING + stage * CHUNK_BYTES);
        const uint64_t wl = make_kmajor_sw128_desc(sbase + OFF_RING + (stage + 1) * CHUNK_BYTES);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, xh + 2 * k, wh + 2 * k, idesc, (kc | k) ? 1u : 0u);
        }
        if (first_of_tile && kc == 0) {                     // the lo half of X arrives later (see xl_full)
          prof_add<PROF>(w_xf, mbar_wait(xl_full, it & 1));
          tc_fence_after();
        }
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, xl + 2 * k, wh + 2 * k, idesc, 1u);
          umma_commit(r_empty(stage));                      // the hi tile goes back a third of a group early: the ring is the bottleneck (r4r)
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, xh + 2 * k, wl + 2 * k, idesc, 1u);
          umma_commit(r_empty(stage + 1));
          if (kc == NKC - 1) umma_commit(s_full(buf));
        }
        __syncwarp();
        stage += 2;
        if (stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    auto mma2 = [&](uint32_t T, bool first, bool last) {   // OUT += H(T) W2(T)^T, H read from TMEM
      prof_add<PROF>(w_hf, mbar_wait(h_full, T & 1));
      if (first) prof_add<PROF>(w_oe, mbar_wait(o_empty, (it & 1) ^ 1));          // the LayerNorm epilogue of the previous row tile has drained OUT
      // per 64-wide K half of the block: (W2.hi tile, W2.lo tile) = one ring pair, 12 UMMAs -- the same bytes per UMMA cycle as a
      // UMMA-1 group (r5: as [hi k0, hi k1 | 16 UMMAs][lo k0, lo k1 | 8 UMMAs] the second pair had to arrive within 512 cycles)
      for (int kh = 0; kh < 2; ++kh) {
        const int s0 = stage;
        prof_add<PROF>(w_rf2, mbar_wait(r_full(s0), phase));
        prof_add<PROF>(w_rf2, mbar_wait(r_full(s0 + 1), phase));
        tc_fence_after();
        if (leader) {
          const uint64_t bh = make_kmajor_sw128_desc(sbase + OFF_RING + s0 * CHUNK_BYTES);
          const uint64_t bl = make_kmajor_sw128_desc(sbase + OFF_RING + (s0 + 1) * CHUNK_BYTES);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t ka = (uint32_t)(8 * (kh * 4 + k4));            // 16 k = 8 packed columns of H
            umma_bf16_ts(t_o, t_h + ka, bh + 2 * k4, idesc, (first && kh == 0 && k4 == 0) ? 0u : 1u);     // H.hi W2.hi
            umma_bf16_ts(t_o, t_h + 64 + ka, bh + 2 * k4, idesc, 1u);                                    // H.lo W2.hi
          }
          umma_commit(r_empty(s0));
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) umma_bf16_ts(t_o, t_h + (uint32_t)(8 * (kh * 4 + k4)), bl + 2 * k4, idesc, 1u);   // H.hi W2.lo
          umma_commit(r_empty(s0 + 1));
          if (kh == 1) {
            umma_commit(h_empty);
            if (last) umma_commit(o_full);
          }
        }
        __syncwarp();
        stage += 2;
        if (stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      prof_add<PROF>(w_xf, mbar_wait(x_full, it & 1));
      tc_fence_after();
      // X is read by UMMA-1 only: it is handed back once the LAST S block of the row tile has been issued, so the producer
      // brings in the next row tile under the last two UMMA-2 groups and the LayerNorm epilogue
      mma1(tile, true);
      if (nj == 1 && leader) umma_commit(x_empty);
      __syncwarp();
      for (int j = 0; j < nj; ++j) {
        if (j + 1 < nj) {
          mma1(tile + j + 1, false);
          if (j + 2 == nj && leader) umma_commit(x_empty);
          __syncwarp();
        }
        mma2(tile + j, j == 0, j == nj - 1);
      }
      tile += nj;
    }
    if (PROF && mp.prof && leader) {
      unsigned long long* o = mp.prof + blockIdx.x * 16;
      o[7] = w_se; o[8] = w_rf + w_rf2; o[9] = w_hf; o[10] = w_oe; o[11] = w_xf; o[12] = clock64() - t_begin; o[14] = w_rf2;
    }
    __syncwarp();
  } else {
    // ===================== GELU + split, LayerNorm epilogue =====================
    const int quarter = warp & 3, part = warp >> 2;        // TMEM lane quarter; 32-column part of every 128-column block
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int cb = part * 32;
    uint32_t tile = 0, it = 0;
    long long w_sf = 0, t_g = 0, w_he = 0, t_st = 0, w_of = 0, t_ln = 0;
    const long long t_begin = PROF ? clock64() : 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      for (int j = 0; j < nj; ++j, ++tile) {
        const int buf = tile & 1;
        prof_add<PROF>(w_sf, mbar_wait(s_full(buf), (tile >> 1) & 1));
        long long c0 = PROF ? clock64() : 0;
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32_async(tmem_base + lane_base + (uint32_t)(buf * TN + cb), r);
        tmem_wait(r);
        tc_fence_before();
        mbar_arrive(s_empty(buf));                          // my part of the S block is in registers
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float v0 = __uint_as_float(r[2 * q]), v1 = __uint_as_float(r[2 * q + 1]);
          gelu_epi2(v0, v1);
          const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
          hi[q] = *reinterpret_cast<const uint32_t*>(&h);
          const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - __uint_as_float(hi[q] << 16), v1 - __uint_as_float(hi[q] & 0xffff0000u));
          lo[q] = *reinterpret_cast<const uint32_t*>(&l);
        }
        // the H buffer is free once UMMA-2 of the previous block has retired
        if (PROF) { const long long c1 = clock64(); t_g += c1 - c0; }
        prof_add<PROF>(w_he, mbar_wait(h_empty, (tile & 1) ^ 1));
        if (PROF) c0 = clock64();
        tc_fence_after();
        const uint32_t t_hh = tmem_base + lane_base + COL_H + (uint32_t)(part * 16);
        tmem_st16(t_hh, hi);
        tmem_st16(t_hh + 64, lo);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(h_full);
        if (PROF) t_st += clock64() - c0;
      }
      // ---- LayerNorm over the 128 outputs of a row (+ residual).  The residual segments this lane will add after the transpose
      // below are requested first (eight independent 16-byte loads in flight under the wait for OUT and the statistics).
      const int cc = cb + (lane & 7) * 4;
      float4 q8[8];
#pragma unroll
      for (int it2 = 0; it2 < 8; ++it2) {
        const int grow = item * TM + quarter * 32 + it2 * 4 + (lane >> 3);
        q8[it2] = (p.res != nullptr && grow < p.L) ? __ldg(reinterpret_cast<const float4*>(p.res + (size_t)grow * p.ldr + cc)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + cc));
      const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + cc));
      prof_add<PROF>(w_of, mbar_wait(o_full, it & 1));
      const long long c_ln = PROF ? clock64() : 0;
      tc_fence_after();
      uint32_t mine[32];
      tmem_ld32_async(tmem_base + lane_base + COL_O + (uint32_t)cb, mine);
      tmem_wait(mine);
      tc_fence_before();
      mbar_arrive(o_empty);                                 // my part of OUT is in registers: the next row tile may accumulate
      // statistics: every thread sums its own 32 columns around their first element; the four parts of a row (four warps of one
      // lane quarter) exchange (mean, sum of squared deviations) through shared memory and merge them (Chan et al.) --
      // one TMEM read per thread instead of the whole row (r4q: 5 x 4 KB per warp and row tile)
      const float cshift = __uint_as_float(mine[0]);
      float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float d = __uint_as_float(mine[i]) - cshift;
        a4[i & 3] += d;
        b4[i & 3] = fmaf(d, d, b4[i & 3]);
      }
      const float s1 = (a4[0] + a4[1]) + (a4[2] + a4[3]), s2 = (b4[0] + b4[1]) + (b4[2] + b4[3]);
      const float m1 = s1 * (1.f / 32.f);
      const uint32_t s_x0 = sbase + OFF_X + NKC * CHUNK_BYTES;                // the staging tiles (lo half of X): 16 warps x 4 KB
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(s_x0 + warp * 4096 + lane * 8), "f"(cshift + m1), "f"(fmaxf(s2 - s1 * m1, 0.f)) : "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
      float pm[4], pq[4];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp)
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(pm[pp]), "=f"(pq[pp]) : "r"(s_x0 + (pp * 4 + quarter) * 4096 + lane * 8) : "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");       // all four parts have read: the tiles may be overwritten
      const float mean = ((pm[0] + pm[1]) + (pm[2] + pm[3])) * 0.25f;
      float dev = 0.f;
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) dev = fmaf(pm[pp] - mean, pm[pp] - mean, dev);
      const float sq = ((pq[0] + pq[1]) + (pq[2] + pq[3])) + 32.f * dev;
      const float rstd = rsqrtf(sq * (1.f / 128.f) + p.eps);
      // my 32 normalised columns go through the warp's staging tile (the lo half of X, free until the next row tile's lo half is
      // loaded; tc_common.cuh::stg_swz) and come out as 4 rows x 128 B per warp instruction: residual load, fp32 store and the
      // bf16 hi | lo stores are then full row segments (r4p: as row-per-thread accesses -- 32 lines per instruction -- this
      // epilogue took 23 k cycles per row tile, a third of the kernel)
      const uint32_t s_stg = sbase + OFF_X + NKC * CHUNK_BYTES + warp * 4096;
      const uint32_t wr_base = s_stg + lane * 128, wr_key = stg_swz(lane);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        sts128(wr_base + ((c ^ wr_key) << 4), make_float4((__uint_as_float(mine[4 * c]) - mean) * rstd, (__uint_as_float(mine[4 * c + 1]) - mean) * rstd,
                                                          (__uint_as_float(mine[4 * c + 2]) - mean) * rstd, (__uint_as_float(mine[4 * c + 3]) - mean) * rstd));
      __syncwarp();
      float4 v8[8];
#pragma unroll
      for (int it2 = 0; it2 < 8; ++it2) {
        const uint32_t rr = it2 * 4 + (lane >> 3);
        v8[it2] = lds128(s_stg + rr * 128 + (((lane & 7) ^ stg_swz(rr)) << 4));
      }
      mbar_arrive(stg_free);                                // my reads of the staging tile are done (ld.shared results are in registers)
#pragma unroll
      for (int it2 = 0; it2 < 8; ++it2) {
        const int grow = item * TM + quarter * 32 + it2 * 4 + (lane >> 3);
        if (grow < p.L) {
          float4 v = make_float4(v8[it2].x * g.x + be.x, v8[it2].y * g.y + be.y, v8[it2].z * g.z + be.z, v8[it2].w * g.w + be.w);
          v.x += q8[it2].x; v.y += q8[it2].y; v.z += q8[it2].z; v.w += q8[it2].w;
          if (p.y != nullptr) *reinterpret_cast<float4*>(p.y + (size_t)grow * p.ldy + cc) = v;
          if (p.out_hi != nullptr) {
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
            const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __uint_as_float(u0 << 16), v.y - __uint_as_float(u0 & 0xffff0000u));
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __uint_as_float(u1 << 16), v.w - __uint_as_float(u1 & 0xffff0000u));
            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out_hi) + (size_t)grow * p.out_ld + cc) = make_uint2(u0, u1);
            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out_lo) + (size_t)grow * p.out_ld + cc) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
          }
        }
      }
      if (PROF) t_ln += clock64() - c_ln;
    }
    if (PROF && mp.prof && threadIdx.x == 0) {
      unsigned long long* o = mp.prof + blockIdx.x * 16;
      o[0] = w_sf; o[1] = t_g; o[2] = w_he; o[3] = t_st; o[4] = w_of; o[5] = t_ln; o[6] = clock64() - t_begin;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NMATH + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace

bool mlp_fused_supported(const MlpFusedArgs& a) {
  const uintptr_t al16 = reinterpret_cast<uintptr_t>(a.x_hi) | reinterpret_cast<uintptr_t>(a.x_lo) | reinterpret_cast<uintptr_t>(a.w1) |
                         reinterpret_cast<uintptr_t>(a.w2) | reinterpret_cast<uintptr_t>(a.gamma) | reinterpret_cast<uintptr_t>(a.beta) |
                         reinterpret_cast<uintptr_t>(a.res) | reinterpret_cast<uintptr_t>(a.y) | reinterpret_cast<uintptr_t>(a.out_hi) |
                         reinterpret_cast<uintptr_t>(a.out_lo);
  return a.L >= 1 && a.hid >= 128 && a.hid % 128 == 0 && a.hid <= 8192 && a.x_hi && a.x_lo && a.w1 && a.w2 && a.gamma && a.beta &&
         (a.y || a.out_hi) && (!a.out_hi || a.out_lo) && a.ldx >= KX && a.ldx % 8 == 0 && (al16 & 15) == 0 &&
         (!a.res || a.ldr % 4 == 0) && (!a.y || a.ldy % 4 == 0) && (!a.out_hi || a.out_ld % 8 == 0);
}

int mlp_fused_tc(const MlpFusedArgs& a, cudaStream_t st) {
  if (a.L == 0) return EMIP_OK;
  if (!mlp_fused_supported(a)) { emip_set_error("mlp_fused_tc: unsupported arguments (256 -> hid %% 128 == 0 -> 128, 16-byte aligned rows)"); return EMIP_ENOSYS; }
  CUtensorMap mxh, mxl, mw1, mw2;
  int rc;
  if ((rc = make_bf16_map(&mxh, a.x_hi, KX, (uint64_t)a.L, 1, (uint64_t)a.ldx * 2, (uint64_t)a.L * a.ldx * 2))) return rc;
  if ((rc = make_bf16_map(&mxl, a.x_lo, KX, (uint64_t)a.L, 1, (uint64_t)a.ldx * 2, (uint64_t)a.L * a.ldx * 2))) return rc;
  if ((rc = make_bf16_map(&mw1, a.w1, 2 * KX, (uint64_t)a.hid, 1, (uint64_t)2 * KX * 2, (uint64_t)a.hid * 2 * KX * 2))) return rc;
  if ((rc = make_bf16_map(&mw2, a.w2, (uint64_t)2 * a.hid, TM, 1, (uint64_t)2 * a.hid * 2, (uint64_t)TM * 2 * a.hid * 2))) return rc;
  if (int rc__ = emip_func_max_smem((const void*)mlp_fused_kernel<false>, SMEM_BYTES)) return rc__;
  if (int rc__ = emip_func_max_smem((const void*)mlp_fused_kernel<true>, SMEM_BYTES)) return rc__;
  MParams mp;
  mp.prof = attn_tc_profile_buffer();
  mp.a = a;
  mp.n_tiles = (a.L + TM - 1) / TM;
  mp.nj = a.hid / TN;
  const int grid = mp.n_tiles < emip_num_sms() ? mp.n_tiles : emip_num_sms();
  if (mp.prof) mlp_fused_kernel<true><<<grid, NTHREADS, SMEM_BYTES, st>>>(mxh, mxl, mw1, mw2, mp);      // diagnostics build
  else EMIP_CUDA(emip_launch_pdl(mlp_fused_kernel<false>, dim3(grid), dim3(NTHREADS), (size_t)SMEM_BYTES, st, mxh, mxl, mw1, mw2, mp));
  EMIP_CHECK_LAUNCH("mlp_fused_tc");
  return EMIP_OK;
}
