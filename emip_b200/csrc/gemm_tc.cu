// Tensor-core GEMMs on bf16 hi|lo split operands: the kernel, the operand re-layout kernels and the gemm_nn / gemm_nt
// front ends (see gemm_tc.cuh).  Used by conv_corr.cu (f1: conv_corr[0] on the never-materialised cost volume) and by
// injector.cu (a4: the 1x1 convolutions of the prompt fusion, forward and backward).
//
// Persistent CTAs (one per SM) walk the 128 x N output tiles: 4 epilogue warps (thread <-> TMEM lane <-> output row),
// 1 TMA producer warp, 1 UMMA issuer warp; operands stream through a 2-3 stage ring of (A.hi, A.lo, B.hi, B.lo) tiles;
// per K chunk of 64 the issuer runs 4 + 4 + 4 UMMAs (hi.hi, lo.hi, hi.lo) into one of two fp32 accumulator tiles in
// TMEM, so the epilogue of a tile overlaps the main loop of the next one.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/emip_b200.h"
#include "gemm_tc.cuh"
#include <math.h>

namespace {
using namespace tc;

constexpr int KCH = GEMM_TC_KCH;
constexpr int A_BYTES = GEMM_TC_A_BYTES;
constexpr int EPI_STAGE_BYTES = 8 * 32 * 128;      // eight epilogue warps x [32 rows][128 B] (swizzled: tc_common.cuh::stg_swz)
constexpr int THREADS = 10 * 32;     // warps 0-3, 6-9: epilogue groups; 4: TMA producer; 5: UMMA issuer

// EPI selects the epilogue at compile time (0: mode 0 | 1: mode 1 | 2: mode 2 + LayerNorm | 3: mode 2 bf16 hi | lo output |
// 4: mode 2 plain fp32): one function for all of them put every branch under the register allocation of the hungriest one
// (168 registers at 10 warps) and spilled; per-instantiation allocation has no spills and lets the split epilogue keep two
// TMEM chunks in flight
template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                      const __grid_constant__ CUtensorMap map_b, const GemmTcParams p) {
  pdl_trigger();                                          // the next kernel's prologue may overlap this kernel's tail (common.cuh)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + p.stages * p.stage_bytes;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 8 * (p.stages + s); };
  // two accumulator tiles of <= 256 TMEM columns: the epilogue of tile i overlaps the main loop of tile i + 1
  // accumulator tiles in TMEM: two of <= 256 columns, or FOUR of <= 128 columns -- then each epilogue group owns two and the
  // MMAs of its next tile run while it still drains the previous one (r3b role profile: with one accumulator per group the
  // issuer waited 43 % of the 256 -> 1024 launch for `acc_empty` and the epilogue warps 31 % for `acc_full`: MMA and epilogue
  // of a group ran strictly one after the other)
  const int nacc = p.n_tile <= 128 ? 4 : 2, acc_cols = 512 / nacc, acc_shift = nacc == 4 ? 2 : 1;
  auto acc_full = [&](int a) { return bar0 + 8 * (2 * p.stages + a); };
  auto acc_empty = [&](int a) { return bar0 + 8 * (2 * p.stages + 4 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.stages * p.stage_bytes + 8 * (2 * p.stages + 8));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ... of the (batch, m tile, n tile) list
  auto decode = [&](int t, int& nt, int& mt, int& b) {
    nt = t % p.n_ntiles; t /= p.n_ntiles;
    mt = t % p.n_mtiles;
    b = t / p.n_mtiles;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 4; ++a) { mbar_init(acc_full(a), 1); mbar_init(acc_empty(a), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // EPI 2: gamma | beta (2 x 128 floats) behind the staging tiles -- read once per 32-column chunk by every epilogue warp (r4d:
  // as __ldg in the chunk loop the pair sat on the long scoreboard in front of the first FFMA, ~20 % of the epilogue warps' samples)
  const uint32_t s_gb = sbase + ((p.stages * p.stage_bytes + 8 * (2 * p.stages + 8) + 16 + 15) & ~15) + EPI_STAGE_BYTES;
  if constexpr (EPI == 2) {
    if (threadIdx.x < 256) {
      const float v = threadIdx.x < 128 ? __ldg(p.ln_gamma + threadIdx.x) : __ldg(p.ln_beta + threadIdx.x - 128);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_gb + threadIdx.x * 4), "f"(v) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                             // everything below reads what the kernel in front of this one wrote

  if (warp == 4) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    long long w_empty = 0;
    const long long t_begin = clock64();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    int nt, mt, b;
    decode(tile, nt, mt, b);
    for (int kc = 0; kc < p.kchunks; ++kc) {
      w_empty += mbar_wait(empty(stage), phase ^ 1);
      if (leader && (p.dbg & 2)) {
        mbar_arrive(full(stage));                              // diagnostics (bit 3): no operand loads at all -- UMMA issue rate alone
      } else if (leader) {
        const uint32_t sa = sbase + stage * p.stage_bytes;
        mbar_expect_tx(full(stage), 2 * A_BYTES + 2 * p.b_bytes);
        if (p.mode == 0) {
          tma_load_3d(sa, &map_a_hi, full(stage), kc * KCH, mt * TM, 0);
          tma_load_3d(sa + A_BYTES, &map_a_lo, full(stage), kc * KCH, mt * TM, 0);
          // f1 channel-major [b][256 = hi c | lo c][ld]: rows 0..127 hi, 128..255 lo
          tma_load_3d(sa + 2 * A_BYTES, &map_b, full(stage), kc * KCH, 0, b);
          tma_load_3d(sa + 2 * A_BYTES + p.b_bytes, &map_b, full(stage), kc * KCH, 128, b);
        } else if (p.mode == 2 && p.b_mn) {
          // B as it lies in memory, [k rows][n contiguous] (hi | lo per row, lo at column Np): two {64 n x 64 k} boxes
          // per operand half = the MN-major SW128 layout (tc_common.cuh), no token-major copy of the activations
          const int ab = p.a_batched ? b : 0;
          tma_load_3d(sa, &map_a_hi, full(stage), kc * KCH, mt * TM, ab);
          tma_load_3d(sa + A_BYTES, &map_a_lo, full(stage), kc * KCH, mt * TM, ab);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tma_load_3d(sa + 2 * A_BYTES + h * 8192, &map_b, full(stage), nt * TM + h * 64, kc * KCH, b);
            tma_load_3d(sa + 2 * A_BYTES + p.b_bytes + h * 8192, &map_b, full(stage), p.Np + nt * TM + h * 64, kc * KCH, b);
          }
        } else if (p.mode == 2) {
          // ksplit > 1: batch entry b = (sample, K split): the split walks its own kchunks chunks of the contraction
          // axis and writes its own partial (K chunks past the operands are zero-filled by TMA on the A side)
          const int ks = p.ksplit > 1 ? p.ksplit : 1, bs = b / ks, k0 = (b - bs * ks) * p.kchunks + kc;
          const int ab = p.a_batched ? bs : 0;
          tma_load_3d(sa, &map_a_hi, full(stage), k0 * KCH, mt * TM, ab);
          tma_load_3d(sa + A_BYTES, &map_a_lo, full(stage), k0 * KCH, mt * TM, ab);
          tma_load_3d(sa + 2 * A_BYTES, &map_b, full(stage), k0 * KCH, nt * p.n_tile, bs);
          tma_load_3d(sa + 2 * A_BYTES + p.b_bytes, &map_b, full(stage), p.Kp + k0 * KCH, nt * p.n_tile, bs);
        } else {
          const int ab = p.a_shared ? 0 : b;                   // conv_corr: per-sample G; generic conv: one weight
          tma_load_3d(sa, &map_a_hi, full(stage), kc * KCH, mt * TM, ab);
          tma_load_3d(sa + A_BYTES, &map_a_lo, full(stage), kc * KCH, mt * TM, ab);
          // K chunk kc = tap (dy,dx) = kc / cpt, channel chunk kc % cpt of the token-major [.., hi Cp | lo Cp] input
          const int tap = kc / p.cpt, dy = tap / 3, dx = tap - dy * 3, c0 = (kc - tap * p.cpt) * KCH;
          tma_load_4d(sa + 2 * A_BYTES, &map_b, full(stage), c0, dx - 1, nt * p.R + dy - 1, b);
          tma_load_4d(sa + 2 * A_BYTES + p.b_bytes, &map_b, full(stage), p.lo_off + c0, dx - 1, nt * p.R + dy - 1, b);
        }
      }
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    }
    __syncwarp();
    if (p.prof != nullptr && lane == 0) { p.prof[blockIdx.x * 8 + 0] += clock64() - t_begin; p.prof[blockIdx.x * 8 + 1] += w_empty; }
  } else if (warp == 5) {
    // ===================== UMMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(p.n_tile);
    int stage = 0;
    uint32_t phase = 0, it = 0;
    long long w_full = 0, w_acc = 0;
    const long long t_begin = clock64();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
    const int acc = it & (nacc - 1);
    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols);
    w_acc += mbar_wait(acc_empty(acc), ((it >> acc_shift) & 1) ^ 1);       // the epilogue has drained this accumulator (nacc tiles ago)
    tc_fence_after();
    for (int kc = 0; kc < p.kchunks; ++kc) {
      w_full += mbar_wait(full(stage), phase);
      tc_fence_after();
      if (leader) {
        const uint32_t sa = sbase + stage * p.stage_bytes;
        const uint64_t a_hi = make_kmajor_sw128_desc(sa), a_lo = make_kmajor_sw128_desc(sa + A_BYTES);
        if (p.b_mn) {
          // MN-major B: a K16 step = two 8-row groups = 2 KB down the box; the second 64 columns 8 KB further (LBO)
          const uint64_t b_hi = make_mnmajor_sw128_desc(sa + 2 * A_BYTES, 8192), b_lo = make_mnmajor_sw128_desc(sa + 2 * A_BYTES + p.b_bytes, 8192);
          const uint32_t idesc_mn = idesc | (1u << 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_hi + 2 * k, b_hi + 128 * k, idesc_mn, (kc | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_lo + 2 * k, b_hi + 128 * k, idesc_mn, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_hi + 2 * k, b_lo + 128 * k, idesc_mn, 1u);
        } else {
        const uint64_t b_hi = make_kmajor_sw128_desc(sa + 2 * A_BYTES), b_lo = make_kmajor_sw128_desc(sa + 2 * A_BYTES + p.b_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (kc | k) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
        }
        umma_commit(empty(stage));
        if (kc == p.kchunks - 1) umma_commit(acc_full(acc));
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    }
    if (p.prof != nullptr && lane == 0) {
      p.prof[blockIdx.x * 8 + 2] += clock64() - t_begin; p.prof[blockIdx.x * 8 + 3] += w_full; p.prof[blockIdx.x * 8 + 4] += w_acc;
    }
  } else {
    // ===================== epilogue warps: thread <-> output row =====================
    // two groups of four warps (0-3 and 6-9), one per accumulator: group g drains the CTA's tiles g, g + 2, ... so two
    // epilogues run side by side (r3: with K <= 256 the epilogue, not the MMA loop, bounds a tile).  A warp reads the TMEM
    // lane quarter warp % 4.
    const int grp = warp >= 6 ? 1 : 0, qw = warp & 3;
    // warp-private staging tile [32 rows][128 B] behind the barriers (16-byte pieces swizzled, tc_common.cuh::stg_swz): a thread's
    // 32 accumulator columns go in as a row, come out as 4 rows x 128 B per warp store (r3h: the row-per-thread stores touched
    // 32 lines per instruction)
    uint32_t* stg = reinterpret_cast<uint32_t*>(smem + ((p.stages * p.stage_bytes + 8 * (2 * p.stages + 8) + 16 + 15) & ~15)) +
                    (grp * 4 + qw) * (32 * 32);
    uint32_t it = grp;
    long long w_accf = 0, n_tiles = 0;
    const long long t_begin = clock64();
    for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < p.total_tiles; tile += 2 * (int)gridDim.x, it += 2) {
    int nt, mt, b;
    decode(tile, nt, mt, b);
    const int acc = it & (nacc - 1);
    const int row = mt * TM + qw * 32 + lane;
    w_accf += mbar_wait(acc_full(acc), (it >> acc_shift) & 1);
    ++n_tiles;
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)(acc * acc_cols);
    if (p.dbg & 1) {
      // diagnostics only (emip_debug_gemm_wide_tiles bit 1): the accumulator is handed back unread
    } else if constexpr (EPI == 0) {
      __nv_bfloat16* gh = p.g_hi + ((size_t)b * p.M + row) * 128;
      __nv_bfloat16* gl = p.g_lo + ((size_t)b * p.M + row) * 128;
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        tmem_ld32_async(taddr + c0, r);
        tmem_wait(r);
        if (row < p.M) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float v0 = __uint_as_float(r[2 * i]) * p.scale, v1 = __uint_as_float(r[2 * i + 1]) * p.scale;
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
            const __nv_bfloat162 h = __halves2bfloat162(h0, h1);
            const __nv_bfloat162 l = __halves2bfloat162(__float2bfloat16_rn(v0 - __bfloat162float(h0)),
                                                         __float2bfloat16_rn(v1 - __bfloat162float(h1)));
            hi[i] = *reinterpret_cast<const uint32_t*>(&h);
            lo[i] = *reinterpret_cast<const uint32_t*>(&l);
          }
          uint4* dh = reinterpret_cast<uint4*>(gh + c0);
          uint4* dl = reinterpret_cast<uint4*>(gl + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            dh[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
            dl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
          }
        }
      }
    } else if constexpr (EPI == 2) {
      // the whole 128-wide output row lives in this thread's TMEM lane: LayerNorm over it (+ residual) before the store --
      // three passes over TMEM (mean, centred variance, write), no fp32 round trip of the pre-norm tensor through HBM
      // residual rows of the staged third pass: all eight 16-byte loads of a chunk are issued together and one chunk ahead
      // (r2k role profile: one dependent load per row segment made the epilogue 38.7 k cycles per tile -- eight HBM
      // round trips per chunk in sequence -- and the whole GEMM epilogue-bound)
      const int ecc = (lane & 7) * 4;
      const uint32_t s_stg = smem_u32(stg), wr_base = s_stg + lane * 128, wr_key = stg_swz(lane);      // staging: see EPI 3
      float4 qn[8];
      auto load_res = [&](int c0) {
#pragma unroll
        for (int it2 = 0; it2 < 8; ++it2) {
          const int grow = mt * TM + qw * 32 + it2 * 4 + (lane >> 3);
          qn[it2] = (p.res != nullptr && grow < p.M) ? __ldg(reinterpret_cast<const float4*>(p.res + (size_t)grow * p.ldr + c0 + ecc))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_res(0);
      // statistics in ONE pass over the TMEM lane (r3c: three passes made this epilogue 18 k cycles per tile, the bound of every
      // merge + norm1 launch): sums of (x - c) and (x - c)^2 around c = the row's first element, so that the variance
      // s2/n - (s1/n)^2 cancels at most one bit (|c - mean| is of the order of the row's own deviation)
      float s1 = 0.f, s2 = 0.f, cshift = 0.f;
      for (int c0 = 0; c0 < TM; c0 += 32) {
        uint32_t r[32];
        tmem_ld32_async(taddr + c0, r);
        tmem_wait(r);
        if (c0 == 0) cshift = __uint_as_float(r[0]);
        float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = __uint_as_float(r[i]) - cshift;
          a4[i & 3] += d;
          b4[i & 3] = fmaf(d, d, b4[i & 3]);
        }
        s1 += (a4[0] + a4[1]) + (a4[2] + a4[3]);
        s2 += (b4[0] + b4[1]) + (b4[2] + b4[3]);
      }
      const float m1 = s1 * (1.f / 128.f);
      const float mean = cshift + m1;
      const float sq = fmaxf(s2 - s1 * m1, 0.f);                   // = sum (x - mean)^2
      const float rstd = rsqrtf(sq * (1.f / 128.f) + p.ln_eps);
      // second pass: the normalised chunk goes through the warp-private staging tile so that the residual read, the fp32 store
      // and the bf16 hi | lo stores are row segments of 128 / 64 bytes (r2: one 16-byte piece per row and instruction made
      // merge + norm1 224 us instead of 108 at 64 pairs)
      for (int c0 = 0; c0 < TM; c0 += 32) {
        uint32_t r[32];
        tmem_ld32_async(taddr + c0, r);
        tmem_wait(r);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(wr_base + ((c ^ wr_key) << 4), make_float4((__uint_as_float(r[4 * c]) - mean) * rstd, (__uint_as_float(r[4 * c + 1]) - mean) * rstd,
                                                            (__uint_as_float(r[4 * c + 2]) - mean) * rstd, (__uint_as_float(r[4 * c + 3]) - mean) * rstd));
        __syncwarp();
        float4 v8[8];
#pragma unroll
        for (int it2 = 0; it2 < 8; ++it2) {
          const uint32_t rr = it2 * 4 + (lane >> 3);
          v8[it2] = lds128(s_stg + rr * 128 + (((lane & 7) ^ stg_swz(rr)) << 4));
        }
        const int cc = ecc;
        const float4 g = lds128(s_gb + (c0 + cc) * 4), be = lds128(s_gb + 512 + (c0 + cc) * 4);
        float4 qc[8];
#pragma unroll
        for (int it2 = 0; it2 < 8; ++it2) qc[it2] = qn[it2];
        if (c0 + 32 < TM) load_res(c0 + 32);
#pragma unroll
        for (int it2 = 0; it2 < 8; ++it2) {
          const int rr = it2 * 4 + (lane >> 3);
          const int grow = mt * TM + qw * 32 + rr;
          float4 v = make_float4(v8[it2].x * g.x + be.x, v8[it2].y * g.y + be.y, v8[it2].z * g.z + be.z, v8[it2].w * g.w + be.w);
          if (grow < p.M) {
            v.x += qc[it2].x; v.y += qc[it2].y; v.z += qc[it2].z; v.w += qc[it2].w;
            if (p.y != nullptr) *reinterpret_cast<float4*>(p.y + (size_t)grow * p.ldy + c0 + cc) = v;
            if (p.ln_hi != nullptr) {                            // the same row as the next GEMM's pre-split A operand
              const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
              const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
              const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __uint_as_float(u0 << 16), v.y - __uint_as_float(u0 & 0xffff0000u));
              const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __uint_as_float(u1 << 16), v.w - __uint_as_float(u1 & 0xffff0000u));
              *reinterpret_cast<uint2*>(p.ln_hi + (size_t)grow * p.ln_ld + c0 + cc) = make_uint2(u0, u1);
              *reinterpret_cast<uint2*>(p.ln_lo + (size_t)grow * p.ln_ld + c0 + cc) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
            }
          }
        }
        __syncwarp();
      }
    } else if constexpr (EPI == 3) {
      // the tile as the bf16 hi | lo A operand of the next GEMM (optionally through the exact GELU): the fp32 tensor never exists
      const int n0 = nt * p.n_tile;
      // The four rows this lane stores per chunk (rows it2 * 8 + lane / 4 of the warp's 32, 16 bytes = 8 columns each): their
      // addresses are the same for every chunk -- one computation (out_split == 3: one map lookup) per tile
      __nv_bfloat16* wrow[4] = {nullptr, nullptr, nullptr, nullptr};
      ptrdiff_t lo_delta = 128;
#pragma unroll
      for (int it2 = 0; it2 < 4; ++it2) {
        const int grow = mt * TM + qw * 32 + it2 * 8 + (lane >> 2);
        if (grow < p.M) {
          if (p.out_split == 3) {                         // row (image, pixel) -> its slot in the window-ordered operand
            const int img = grow / p.npix, pix = grow - img * p.npix;
            const int2 m = __ldg(p.rowmap + pix);
            int im2 = img + p.img_shift;
            if (im2 >= p.n_img) im2 -= p.n_img;
            __nv_bfloat16* dbase = nt == 0 ? p.split_dst[0] : nt == 1 ? p.split_dst[1] : nt == 2 ? p.split_dst[2] : p.split_dst[3];
            wrow[it2] = dbase + ((size_t)m.x + (size_t)im2 * m.y) * 256 + (lane & 3) * 8;
          } else {
            wrow[it2] = p.g_hi + (size_t)grow * p.ldg + n0 + (lane & 3) * 8;
          }
        }
      }
      if (p.out_split != 3) lo_delta = p.g_lo - p.g_hi;
      // staging tile (tc_common.cuh::stg_swz): this lane's row goes in as eight 16-byte pieces (hi 0..3 | lo 4..7), comes out as
      // piece lane % 4 of hi and of lo of the rows above: 8 STS.128 + 8 LDS.128 per chunk, no bank conflicts (r4a: the [32][33]
      // word tile with scalar generic LD / ST was 50 % of the epilogue warps' samples -- 64 instructions per chunk, the reads 2-way
      // conflicted, every store waiting on the long scoreboard)
      const uint32_t s_stg = smem_u32(stg);
      const uint32_t wr_base = s_stg + lane * 128, wr_key = stg_swz(lane), rd_key = stg_swz(lane >> 2);
      const uint32_t rd_hi = s_stg + (lane >> 2) * 128 + (((lane & 3) ^ rd_key) << 4);
      const uint32_t rd_lo = s_stg + (lane >> 2) * 128 + ((((lane & 3) + 4) ^ rd_key) << 4);
      // one 32-column chunk: r = the accumulator chunk just read from TMEM
      auto chunk = [&](uint32_t (&r)[32], const int c0) {
        if (n0 + c0 >= p.N) return;                            // warp-uniform
        if (row < p.M) {
          uint32_t hi[16], lo[16];
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (p.bias_n != nullptr) {                          // per-column bias (nn.Linear with bias feeding another GEMM), uniform branch
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bq = __ldg(reinterpret_cast<const float4*>(p.bias_n + n0 + c0) + i);
              v[4 * i] += bq.x; v[4 * i + 1] += bq.y; v[4 * i + 2] += bq.z; v[4 * i + 3] += bq.w;
            }
          }
          if (p.out_split == 2) {                             // one uniform branch per chunk, not per element
#pragma unroll
            for (int i = 0; i < 16; ++i) gelu_epi2(v[2 * i], v[2 * i + 1]);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);            // one packed conversion
            hi[i] = *reinterpret_cast<const uint32_t*>(&h);
            const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * i] - __uint_as_float(hi[i] << 16),
                                                           v[2 * i + 1] - __uint_as_float(hi[i] & 0xffff0000u));
            lo[i] = *reinterpret_cast<const uint32_t*>(&l);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            sts128u(wr_base + ((c ^ wr_key) << 4), hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
            sts128u(wr_base + (((c + 4) ^ wr_key) << 4), lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
          }
        }
        __syncwarp();
        uint4 vh[4], vl[4];
#pragma unroll
        for (int it2 = 0; it2 < 4; ++it2) { vh[it2] = lds128u(rd_hi + it2 * 1024); vl[it2] = lds128u(rd_lo + it2 * 1024); }
#pragma unroll
        for (int it2 = 0; it2 < 4; ++it2) {
          if (wrow[it2] != nullptr) {
            *reinterpret_cast<uint4*>(wrow[it2] + c0) = vh[it2];
            *reinterpret_cast<uint4*>(wrow[it2] + c0 + lo_delta) = vl[it2];
          }
        }
        __syncwarp();
      };
      // (r2u: keeping the next chunk's TMEM read in flight while this one is converted did not help -- the GELU / split
      // epilogue is issue- and dependency-bound, not TMEM-latency-bound -- and cost the 256 -> 1024 layer 10 %: not kept)
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        uint32_t r[32];
        tmem_ld32_async(taddr + c0, r);
        tmem_wait(r);
        chunk(r, c0);
      }
    } else if constexpr (EPI == 4) {
      const int n0 = nt * p.n_tile;
      float* o = p.y + (size_t)b * p.y_stride_b + (size_t)row * p.ldy + n0;
      const float* rs = p.res ? p.res + (size_t)b * p.res_stride_b + (size_t)row * p.ldr + n0 : nullptr;
      const float rb = (p.bias != nullptr && row < p.M) ? __ldg(p.bias + row) : 0.f;      // per-row bias (linear layers)
      const uint32_t s_stg = smem_u32(stg), wr_base = s_stg + lane * 128, wr_key = stg_swz(lane);      // staging: see EPI 3
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        const int n = min(32, p.N - n0 - c0);
        const float* yb = p.y + (size_t)b * p.y_stride_b + n0 + c0;
        const float* rb0 = p.res ? p.res + (size_t)b * p.res_stride_b + n0 + c0 : nullptr;
        const bool staged = p.epi_stage && n == 32 && (p.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(yb) & 15) == 0 &&
                            (!rb0 || ((p.ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(rb0) & 15) == 0));      // warp-uniform
        // the residual row segments of this chunk: eight independent loads in flight while the accumulator chunk is read from
        // TMEM and staged (one dependent load per store made the residual GEMMs of the prompt fusion 3x slower than the plain ones)
        float4 q8[8];
        if (staged && rb0 != nullptr) {
#pragma unroll
          for (int it2 = 0; it2 < 8; ++it2) {
            const int grow = mt * TM + qw * 32 + it2 * 4 + (lane >> 3);
            q8[it2] = grow < p.M ? __ldg(reinterpret_cast<const float4*>(rb0 + (size_t)grow * p.ldr + (lane & 7) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        uint32_t r[32];
        tmem_ld32_async(taddr + c0, r);
        tmem_wait(r);
        if (staged) {
          if (p.bias_n != nullptr) {                         // per-column bias (nn.Linear on token rows), warp-uniform branch
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __ldg(p.bias_n + n0 + c0 + i));
          }
#pragma unroll
          for (int c = 0; c < 8; ++c)
            sts128(wr_base + ((c ^ wr_key) << 4), make_float4(__uint_as_float(r[4 * c]) + rb, __uint_as_float(r[4 * c + 1]) + rb,
                                                              __uint_as_float(r[4 * c + 2]) + rb, __uint_as_float(r[4 * c + 3]) + rb));
          __syncwarp();
          float4 v8[8];
#pragma unroll
          for (int it2 = 0; it2 < 8; ++it2) {
            const uint32_t rr = it2 * 4 + (lane >> 3);
            v8[it2] = lds128(s_stg + rr * 128 + (((lane & 7) ^ stg_swz(rr)) << 4));
          }
#pragma unroll
          for (int it2 = 0; it2 < 8; ++it2) {
            const int rr = it2 * 4 + (lane >> 3), cc = (lane & 7) * 4;
            const int grow = mt * TM + qw * 32 + rr;
            float4 v = v8[it2];
            if (grow < p.M) {
              if (rb0) { v.x += q8[it2].x; v.y += q8[it2].y; v.z += q8[it2].z; v.w += q8[it2].w; }
              *reinterpret_cast<float4*>(const_cast<float*>(yb) + (size_t)grow * p.ldy + cc) = v;
            }
          }
          __syncwarp();
        } else if (row < p.M && n > 0) {
          if (p.bias_n != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < n) r[i] = __float_as_uint(__uint_as_float(r[i]) + __ldg(p.bias_n + n0 + c0 + i));
          }
          if (n == 32 && ((reinterpret_cast<uintptr_t>(o + c0) & 15) == 0) && (!rs || (reinterpret_cast<uintptr_t>(rs + c0) & 15) == 0)) {
            float4* d = reinterpret_cast<float4*>(o + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 v = make_float4(__uint_as_float(r[4 * i]) + rb, __uint_as_float(r[4 * i + 1]) + rb, __uint_as_float(r[4 * i + 2]) + rb,
                                     __uint_as_float(r[4 * i + 3]) + rb);
              if (rs) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(rs + c0) + i);
                v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
              }
              d[i] = v;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < n) o[c0 + i] = __uint_as_float(r[i]) + rb + (rs ? __ldg(rs + c0 + i) : 0.f);
          }
        }
      }
    } else {
      const int npix = p.H * p.W;
      const int px0 = nt * p.R * p.W;
      const int nvalid = min(p.n_tile, npix - px0);          // the last tile may hang over the bottom edge
      // per-output-channel affine + ReLU: bias alone, or a folded eval-mode BatchNorm (scale, shift) of the layer behind
      const float bias = (row < p.M && p.bias != nullptr) ? __ldg(p.bias + row) : 0.f;
      const float esc = (row < p.M && p.ep_scale != nullptr) ? __ldg(p.ep_scale + row) : 1.f;
      const float efl = p.ep_relu ? 0.f : -INFINITY;
      float* o = p.out + ((size_t)b * p.M + row) * npix + px0;
      if (p.tok_hi != nullptr) {
        const uint32_t s_stg = smem_u32(stg);
        // token-major bf16 hi | lo output: transpose each [32 channels x 32 pixels] chunk through the warp's staging tile
        for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
          uint32_t r[32];
          if (c0 + 32 <= p.n_tile) {
            tmem_ld32_async(taddr + c0, r);
            tmem_wait(r);
          } else {
            uint32_t h[32];
            tmem_ld32_async(taddr + p.n_tile - 32, h);
            tmem_wait(h);
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = h[16 + i];
#pragma unroll
            for (int i = 16; i < 32; ++i) r[i] = 0u;
          }
          // staging tile rows = pixels (tc_common.cuh::stg_swz): lane = channel writes word `lane` of every pixel row
#pragma unroll
          for (int i = 0; i < 32; ++i)
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_stg + i * 128 + ((((uint32_t)lane >> 2) ^ stg_swz(i)) << 4) + (lane & 3) * 4),
                         "r"(row < p.M ? __float_as_uint(fmaxf(fmaf(__uint_as_float(r[i]), esc, bias), efl)) : 0u)
                         : "memory");
          __syncwarp();
          const int cc = (lane & 7) * 4;
          float4 v8[8];
#pragma unroll
          for (int it2 = 0; it2 < 8; ++it2) {
            const uint32_t pr = it2 * 4 + (lane >> 3);
            v8[it2] = lds128(s_stg + pr * 128 + (((lane & 7) ^ stg_swz(pr)) << 4));
          }
#pragma unroll
          for (int it2 = 0; it2 < 8; ++it2) {
            const int pr = it2 * 4 + (lane >> 3);                      // pixel of the chunk
            const float v0 = v8[it2].x, v1 = v8[it2].y, v2 = v8[it2].z, v3 = v8[it2].w;
            if (c0 + pr < min(p.n_tile, nvalid)) {
              const __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
              const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
              const __nv_bfloat162 l0 = __floats2bfloat162_rn(v0 - __uint_as_float(u0 << 16), v1 - __uint_as_float(u0 & 0xffff0000u));
              const __nv_bfloat162 l1 = __floats2bfloat162_rn(v2 - __uint_as_float(u1 << 16), v3 - __uint_as_float(u1 & 0xffff0000u));
              __nv_bfloat16* d = p.tok_hi + ((size_t)b * npix + px0 + c0 + pr) * p.tok_ld + mt * TM + qw * 32 + cc;
              *reinterpret_cast<uint2*>(d) = make_uint2(u0, u1);
              *reinterpret_cast<uint2*>(d + p.tok_lo_off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
            }
          }
          __syncwarp();
        }
      } else
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        uint32_t r[32];
        // n_tile is a multiple of 16: the last group may be a half group
        if (c0 + 32 <= p.n_tile) {
          tmem_ld32_async(taddr + c0, r);
          tmem_wait(r);
        } else {
          uint32_t h[32];
          tmem_ld32_async(taddr + p.n_tile - 32, h);          // overlapping read of the last 32 columns
          tmem_wait(h);
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = h[16 + i];
#pragma unroll
          for (int i = 16; i < 32; ++i) r[i] = 0u;
        }
        if (row < p.M) {
          const int n = min(32, min(p.n_tile, nvalid) - c0);
          if (n == 32 && ((reinterpret_cast<uintptr_t>(o + c0) & 15) == 0)) {
            float4* d = reinterpret_cast<float4*>(o + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              d[i] = make_float4(fmaxf(fmaf(__uint_as_float(r[4 * i]), esc, bias), efl), fmaxf(fmaf(__uint_as_float(r[4 * i + 1]), esc, bias), efl),
                                 fmaxf(fmaf(__uint_as_float(r[4 * i + 2]), esc, bias), efl), fmaxf(fmaf(__uint_as_float(r[4 * i + 3]), esc, bias), efl));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < n) o[c0 + i] = fmaxf(fmaf(__uint_as_float(r[i]), esc, bias), efl);
          }
        }
      }
    }
    tc_fence_before();
    mbar_arrive(acc_empty(acc));                          // every TMEM read of this tile has completed
    }
    if (p.prof != nullptr && warp == 0 && lane == 0) {
      p.prof[blockIdx.x * 8 + 5] += clock64() - t_begin; p.prof[blockIdx.x * 8 + 6] += w_accf; p.prof[blockIdx.x * 8 + 7] += n_tiles;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

int make_map_bf16_impl(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box) {
  EncodeTiledFn enc = get_encoder();
  if (enc == nullptr) {
    emip_set_error("cuTensorMapEncodeTiled not available from the driver");
    return EMIP_ENOSYS;
  }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    emip_set_error("gemm_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return EMIP_EINVAL;
  }
  return EMIP_OK;
}


// ---- operand re-layouts: fp32 -> bf16 hi | lo -------------------------------------------------------------------
// 8 fp32 -> 8 bf16 hi + 8 bf16 lo (two 16-byte stores)
__device__ __forceinline__ void split8_store(const float (&v)[8], __nv_bfloat16* hi, __nv_bfloat16* lo) {
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
  *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// weights / row operands W(b)[m][k] fp32 -> hi, lo bf16 [nbw][M][Kp] (zero padded to Kp; Kp % 64 == 0).
// grid (split_w_blocks(M, Kp), 1, nbw): 8 consecutive k per thread, the (row, k group) pairs dealt to threads as one
// flat list (r3: one block per row left 240 of 256 threads idle on the 128-wide token rows -- 36 us for 32 MB).
__global__ void __launch_bounds__(256)
split_w_kernel(const float* __restrict__ w, long long w_stride_b, int ldw, int trans, __nv_bfloat16* __restrict__ hi,
               __nv_bfloat16* __restrict__ lo, int M, int K, int Kp, int act = 0, const float* __restrict__ aux = nullptr) {
  const int b = blockIdx.z;
  const int kg = Kp >> 3;                                                   // k groups per row
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * kg) return;
  const int m = (int)(idx / kg), k0 = (int)(idx % kg) * 8;
  const float* src = w + (size_t)b * w_stride_b;
  float v[8];
  if (!trans && k0 + 8 <= K && (ldw & 3) == 0 && (w_stride_b & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + (size_t)m * ldw + k0));
    const float4 c = __ldg(reinterpret_cast<const float4*>(src + (size_t)m * ldw + k0 + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + i;
      v[i] = k < K ? __ldg(src + (trans ? (size_t)k * ldw + m : (size_t)m * ldw + k)) : 0.f;
    }
  }
  if (act == 1) {                                   // exact GELU (nn.GELU default): 0.5 x (1 + erf(x / sqrt 2))
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.5f * v[i] * (1.f + erff(v[i] * 0.70710678118654752440f));
  }
  if (act == 2) {                                   // v * GELU'(aux), aux laid out like w (not transposed): Phi(h) + h phi(h)
    const float* ax = aux + (size_t)b * w_stride_b + (size_t)m * ldw;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float h = k0 + i < K ? __ldg(ax + k0 + i) : 0.f;
      v[i] *= 0.5f * (1.f + erff(h * 0.70710678118654752440f)) + h * 0.39894228040143267794f * __expf(-0.5f * h * h);
    }
  }
  const size_t o = ((size_t)b * M + m) * Kp + k0;
  split8_store(v, hi + o, lo + o);
}

// B'(b)[k][n] fp32 (rows ldb apart), optional LayerNorm over k -> bf16 [b][k][2*Np] (hi | lo), zero padded to Np.
// grid (ceil(Np / 2048), K, B): 8 consecutive n per thread.
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ x, long long x_stride_b, int ldx, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                  __nv_bfloat16* __restrict__ dst, int K, int N, int Np) {
  const int b = blockIdx.z, k = blockIdx.y;
  const int n0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (n0 >= Np) return;
  const float ga = mean != nullptr ? __ldg(gamma + k) : 1.f, be = mean != nullptr ? __ldg(beta + k) : 0.f;
  const float* src = x + (size_t)b * x_stride_b + (size_t)k * ldx;
  float v[8];
  const bool vec = n0 + 8 <= N && (ldx & 3) == 0 && (x_stride_b & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (N & 3) == 0;
  if (vec) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + n0)), c = __ldg(reinterpret_cast<const float4*>(src + n0 + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = n0 + i < N ? __ldg(src + n0 + i) : 0.f;
  }
  if (mean != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (n0 + i < N) v[i] = (v[i] - __ldg(mean + (size_t)b * N + n0 + i)) * __ldg(rstd + (size_t)b * N + n0 + i) * ga + be;
  }
  __nv_bfloat16* d = dst + ((size_t)b * K + k) * 2 * Np + n0;
  split8_store(v, d, d + Np);
}

int kpad(int K) { return (K + KCH - 1) / KCH * KCH; }
unsigned split_w_blocks(int M, int Kp) { return (unsigned)(((long long)M * (Kp >> 3) + 255) / 256); }

}  // namespace

static int g_gemm_wide_tiles = 0;   // r3c: with four accumulators 128-column tiles are as fast at 64 pairs and 3.5 % faster at 8
// Diagnostics / tuning: 0 (default) = 128-column tiles (four TMEM accumulators) also for wide outputs, 1 = 256-column tiles for N >= 512.
static int g_ft_two_launch_mlp = 0;   // diagnostics: bit 2 -- the FeatureTransformer runs its FFN as two gemm_tc launches instead of mlp_fused.cu
static int g_gemm_dbg = 0;   // diagnostics: bits 1 and 3 of the switch below (tools/gemm_floor.py); results are garbage while set
extern "C" void emip_debug_gemm_wide_tiles(int v) {
  g_gemm_wide_tiles = (v & 1) ? 1 : 0;
  g_gemm_dbg = ((v & 2) ? 1 : 0) | ((v & 8) ? 2 : 0);
  g_ft_two_launch_mlp = (v & 4) ? 1 : 0;
}
int gemm_tc_debug_two_launch_mlp() { return g_ft_two_launch_mlp; }
static unsigned long long* g_gemm_prof = nullptr;
// Diagnostics (tools/gemm_roles.py): device buffer of (SM count) x 8 cycle counters the next launches ADD to; NULL = off.
extern "C" void emip_gemm_tc_set_profile_buffer(unsigned long long* dev_buf) { g_gemm_prof = dev_buf; }

int gemm_tc_make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box) {
  return make_map_bf16_impl(m, base, rank, dims, strides_bytes, box);
}

int gemm_tc_launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b, GemmTcParams p, int batch,
                   cudaStream_t st) {
  const int max_smem = 227 * 1024;
  const int epi = p.mode == 0 ? 0 : p.mode == 1 ? 1 : p.ln_gamma != nullptr ? 2 : p.out_split != 0 ? 3 : 4;
  const void* kfn = epi == 0 ? (const void*)gemm_tc_kernel<0> : epi == 1 ? (const void*)gemm_tc_kernel<1> : epi == 2 ? (const void*)gemm_tc_kernel<2>
                  : epi == 3 ? (const void*)gemm_tc_kernel<3> : (const void*)gemm_tc_kernel<4>;
  if (int rc__ = emip_func_max_smem(kfn, max_smem)) return rc__;
  if (p.stages == 0) {
    p.b_bytes = (int)emip_align_up((size_t)p.n_tile * 128, 1024);
    p.stage_bytes = 2 * A_BYTES + 2 * p.b_bytes;
    p.epi_stage = (p.mode == 2 || (p.mode == 1 && p.tok_hi != nullptr)) ? 1 : 0;
    p.stages = (max_smem - 2048 - (p.epi_stage ? EPI_STAGE_BYTES : 0)) / p.stage_bytes;
    if (p.stages > 3) p.stages = 3;
    // (the ring runs on across tiles: with K = 128 a third stage prefetches the next tile's first chunk)
  }
  p.prof = g_gemm_prof;
  p.dbg = g_gemm_dbg;
  const long long tiles = (long long)batch * p.n_mtiles * p.n_ntiles;
  EMIP_CHECK_ARG(tiles > 0 && tiles < 0x7fffffffLL, "gemm_tc: bad tile count");
  p.total_tiles = (int)tiles;
  const int grid = (int)(tiles < emip_num_sms() ? tiles : emip_num_sms());      // persistent CTAs, one per SM
  const int smem = p.stages * p.stage_bytes + 8 * (2 * p.stages + 8) + 16 + 1024 + (p.epi_stage ? EPI_STAGE_BYTES + 16 : 0) + (epi == 2 ? 1024 : 0);
  switch (epi) {
    case 0: EMIP_CUDA(emip_launch_pdl(gemm_tc_kernel<0>, dim3(grid), dim3(THREADS), (size_t)smem, st, a_hi, a_lo, b, p)); break;
    case 1: EMIP_CUDA(emip_launch_pdl(gemm_tc_kernel<1>, dim3(grid), dim3(THREADS), (size_t)smem, st, a_hi, a_lo, b, p)); break;
    case 2: EMIP_CUDA(emip_launch_pdl(gemm_tc_kernel<2>, dim3(grid), dim3(THREADS), (size_t)smem, st, a_hi, a_lo, b, p)); break;
    case 3: EMIP_CUDA(emip_launch_pdl(gemm_tc_kernel<3>, dim3(grid), dim3(THREADS), (size_t)smem, st, a_hi, a_lo, b, p)); break;
    default: EMIP_CUDA(emip_launch_pdl(gemm_tc_kernel<4>, dim3(grid), dim3(THREADS), (size_t)smem, st, a_hi, a_lo, b, p)); break;
  }
  EMIP_CHECK_LAUNCH("gemm_tc");
  return EMIP_OK;
}

bool gemm_nn_tc_supported(const GemmNN& a) {
  return !a.accumulate && a.K >= 1 && a.K <= 1024 && a.M >= 1 && a.N >= 1 && a.ldy % 4 == 0 &&
         reinterpret_cast<uintptr_t>(a.y) % 16 == 0 && a.y_stride_b % 4 == 0;
}

size_t gemm_nn_tc_scratch_bytes(int B, int M, int K, int N, bool per_sample_w) {
  const size_t Kp = (size_t)kpad(K), Np = (size_t)kpad(N);
  return emip_align_up((size_t)(per_sample_w ? B : 1) * M * Kp * 2 * 2, 1024) + emip_align_up((size_t)B * K * 2 * Np * 2, 1024) + 1024;
}

void* gemm_nn_tc_act_operand(void* scratch, int B, int M, int K, int N, bool per_sample_w, int* np_out) {
  const int Kp = kpad(K), nbw = per_sample_w ? B : 1;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~(uintptr_t)1023);
  if (np_out) *np_out = kpad(N);
  return base + emip_align_up((size_t)nbw * M * Kp * 2 * 2, 1024);
}

int gemm_nn_tc(const GemmNN& a, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  if (a.B == 0 || a.M == 0 || a.N == 0) return EMIP_OK;
  if (!gemm_nn_tc_supported(a)) { emip_set_error("gemm_nn_tc: unsupported arguments"); return EMIP_ENOSYS; }
  const bool per_sample = a.w_stride_b != 0;
  const int Kp = kpad(a.K), nbw = per_sample ? a.B : 1;
  if (scratch == nullptr || scratch_bytes < gemm_nn_tc_scratch_bytes(a.B, a.M, a.K, a.N, per_sample)) {
    emip_set_error("gemm_nn_tc: scratch too small");
    return EMIP_ENOMEM;
  }
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* w_hi = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* w_lo = w_hi + (size_t)nbw * a.M * Kp;
  __nv_bfloat16* bt = reinterpret_cast<__nv_bfloat16*>(base + emip_align_up((size_t)nbw * a.M * Kp * 2 * 2, 1024));
  split_w_kernel<<<dim3(split_w_blocks(a.M, Kp), 1, nbw), 256, 0, st>>>(a.w, a.w_stride_b, a.ldw, a.w_trans, w_hi, w_lo, a.M, a.K, Kp);
  EMIP_CHECK_LAUNCH("gemm_nn_tc (weights)");
  // the activations keep their channel-major layout ([k][n], hi | lo per row): elementwise split (with the LayerNorm
  // applied on the way), read by the GEMM as an MN-major operand.  (r1 / r2a re-laid them out token-major with
  // a transposing kernel: 64-byte store segments, 200 us of the Injector's forward + backward.)
  const int Np = kpad(a.N);
  if (!a.x_presplit) {
  split_rows_kernel<<<dim3((Np + 2047) / 2048, a.K, a.B), 256, 0, st>>>(a.x, a.x_stride_b, a.ldx, a.mean, a.rstd, a.gamma, a.beta, bt,
                                                                    a.K, a.N, Np);
  EMIP_CHECK_LAUNCH("gemm_nn_tc (activations)");
  }
  CUtensorMap ma_hi, ma_lo, mb;
  int rc;
  const cuuint64_t adims[3] = {(cuuint64_t)Kp, (cuuint64_t)a.M, (cuuint64_t)nbw};
  const cuuint64_t astr[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)a.M * Kp * 2};
  const cuuint32_t abox[3] = {KCH, TM, 1};
  if ((rc = gemm_tc_make_map(&ma_hi, w_hi, 3, adims, astr, abox))) return rc;
  if ((rc = gemm_tc_make_map(&ma_lo, w_lo, 3, adims, astr, abox))) return rc;
  const cuuint64_t bdims[3] = {(cuuint64_t)2 * Np, (cuuint64_t)a.K, (cuuint64_t)a.B};
  const cuuint64_t bstr[2] = {(cuuint64_t)2 * Np * 2, (cuuint64_t)a.K * 2 * Np * 2};
  const cuuint32_t bbox[3] = {64, 64, 1};
  if ((rc = gemm_tc_make_map(&mb, bt, 3, bdims, bstr, bbox))) return rc;
  GemmTcParams p = {};
  p.mode = 2; p.M = a.M; p.n_mtiles = (a.M + TM - 1) / TM; p.n_ntiles = (a.N + TM - 1) / TM; p.n_tile = TM;
  p.kchunks = Kp / KCH;
  p.b_mn = 1; p.Np = Np;
  p.bias = a.bias;
  p.a_batched = per_sample ? 1 : 0; p.Kp = Kp; p.N = a.N;
  p.y = a.y; p.y_stride_b = a.y_stride_b; p.ldy = a.ldy;
  p.res = a.res; p.res_stride_b = a.res_stride_b; p.ldr = a.ldr;
  return gemm_tc_launch(ma_hi, ma_lo, mb, p, a.B, st);
}

// C[b][m][k] = sum_n A(b)[m][n] B'(b)[k][n]: the weight-gradient GEMMs (contraction over the contiguous pixel axis, so
// both operands are K-major as they lie in memory and only need the elementwise hi|lo split)
bool gemm_nt_tc_supported(const GemmNT& a) {
  if (a.c_hi != nullptr || a.a_hi_pre != nullptr) {
    if (a.B != 1) return false;
    if (a.c_hi != nullptr && (a.c_lo == nullptr || a.K % 32 != 0 || a.ldc_split % 8 != 0 || a.ldc_split < a.K ||
                              (reinterpret_cast<uintptr_t>(a.c_hi) | reinterpret_cast<uintptr_t>(a.c_lo)) % 16 != 0))
      return false;
    if (a.a_act == 2 && a.a_aux == nullptr) return false;
    if (a.a_hi_pre != nullptr && (a.a_lo_pre == nullptr || a.N % 64 != 0 || a.a_act != 0 ||
                                  (reinterpret_cast<uintptr_t>(a.a_hi_pre) | reinterpret_cast<uintptr_t>(a.a_lo_pre)) % 128 != 0))
      return false;
    if (a.c_hi != nullptr) return a.M >= 1 && a.K >= 1 && a.N >= 16;
  }
  return a.M >= 1 && a.K >= 1 && a.N >= 16 && a.ldc % 4 == 0 && reinterpret_cast<uintptr_t>(a.c) % 16 == 0 && a.c_stride_b % 4 == 0;
}

size_t gemm_nt_tc_scratch_bytes(int B, int M, int K, int N) {
  const size_t Np = (size_t)kpad(N);
  return emip_align_up((size_t)B * M * Np * 2 * 2, 1024) + emip_align_up((size_t)B * K * 2 * Np * 2, 1024) + 1024;
}
// with a pre-split A operand only the B operand needs scratch
// rows [M][K] fp32 -> hi, lo bf16 [M][kpad(K)] (the A operand GemmNT::a_hi_pre takes): one split shared by several GEMMs
int gemm_tc_split_rows(const float* x, int M, int K, void* hi, void* lo, cudaStream_t st) {
  const int Kp = kpad(K);
  split_w_kernel<<<dim3(split_w_blocks(M, Kp), 1, 1), 256, 0, st>>>(x, 0, K, 0, static_cast<__nv_bfloat16*>(hi),
                                                                    static_cast<__nv_bfloat16*>(lo), M, K, Kp);
  EMIP_CHECK_LAUNCH("gemm_tc_split_rows");
  return EMIP_OK;
}
// weight [K][N] fp32 (rows ldb apart) -> the B operand of gemm_nt_tc as GemmNT::b_pre takes it: bf16 [K][2 * kpad(N)] (hi | lo)
int gemm_tc_split_b(const float* w, int ldb, int K, int N, void* dst, cudaStream_t st) {
  const int Np = kpad(N);
  split_rows_kernel<<<dim3((Np + 2047) / 2048, K, 1), 256, 0, st>>>(w, 0, ldb, nullptr, nullptr, nullptr, nullptr,
                                                                 static_cast<__nv_bfloat16*>(dst), K, N, Np);
  EMIP_CHECK_LAUNCH("gemm_tc_split_b");
  return EMIP_OK;
}
size_t gemm_tc_split_b_bytes(int K, int N) { return emip_align_up((size_t)K * 2 * kpad(N) * 2, 1024); }
size_t gemm_nt_tc_scratch_bytes_presplit(int K, int N) { return emip_align_up((size_t)K * 2 * kpad(N) * 2, 1024) + 1024; }

int gemm_nt_tc(const GemmNT& a, void* scratch, size_t scratch_bytes, cudaStream_t st, int nsplit) {
  if (a.B == 0 || a.M == 0 || a.K == 0) return EMIP_OK;
  if (!gemm_nt_tc_supported(a) || (a.c_hi != nullptr && nsplit > 1)) { emip_set_error("gemm_nt_tc: unsupported arguments"); return EMIP_ENOSYS; }
  const bool pre = a.a_hi_pre != nullptr;
  if (a.b_pre == nullptr || !pre) {
  if (scratch == nullptr ||
      scratch_bytes < (pre ? gemm_nt_tc_scratch_bytes_presplit(a.K, a.N) : gemm_nt_tc_scratch_bytes(a.B, a.M, a.K, a.N))) {
    emip_set_error("gemm_nt_tc: scratch too small");
    return EMIP_ENOMEM;
  }
  }
  const int Np = kpad(a.N);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* a_lo = a_hi + (size_t)a.B * a.M * Np;
  __nv_bfloat16* bt = reinterpret_cast<__nv_bfloat16*>(base + (pre ? 0 : emip_align_up((size_t)a.B * a.M * Np * 2 * 2, 1024)));
  if (pre) {
    a_hi = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(a.a_hi_pre));
    a_lo = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(a.a_lo_pre));
  } else {
    split_w_kernel<<<dim3(split_w_blocks(a.M, Np), 1, a.B), 256, 0, st>>>(a.a, a.a_stride_b, a.lda, 0, a_hi, a_lo, a.M, a.N, Np, a.a_act, a.a_aux);
    EMIP_CHECK_LAUNCH("gemm_nt_tc (A)");
  }
  if (a.b_pre != nullptr) {
    bt = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(a.b_pre));
  } else {
  split_rows_kernel<<<dim3((Np + 2047) / 2048, a.K, a.B), 256, 0, st>>>(a.bm, a.b_stride_b, a.ldb, a.mean, a.rstd, a.gamma, a.beta, bt,
                                                                    a.K, a.N, Np);
  EMIP_CHECK_LAUNCH("gemm_nt_tc (B)");
  }
  CUtensorMap ma_hi, ma_lo, mb;
  int rc;
  const cuuint64_t ald = (pre && a.a_ld_pre > 0) ? (cuuint64_t)a.a_ld_pre : (cuuint64_t)Np;
  const cuuint64_t adims[3] = {(cuuint64_t)Np, (cuuint64_t)a.M, (cuuint64_t)a.B};
  const cuuint64_t astr[2] = {ald * 2, (cuuint64_t)a.M * ald * 2};
  const cuuint32_t abox[3] = {KCH, TM, 1};
  if ((rc = gemm_tc_make_map(&ma_hi, a_hi, 3, adims, astr, abox))) return rc;
  if ((rc = gemm_tc_make_map(&ma_lo, a_lo, 3, adims, astr, abox))) return rc;
  const cuuint64_t bdims[3] = {(cuuint64_t)2 * Np, (cuuint64_t)a.K, (cuuint64_t)a.B};
  const cuuint64_t bstr[2] = {(cuuint64_t)2 * Np * 2, (cuuint64_t)a.K * 2 * Np * 2};
  // wide outputs (the 256 -> 1024 MLP layer): 256-column tiles -- the A tile (hi + lo) is fetched once per 256 instead
  // of once per 128 output columns (r3: that GEMM was bound by the L2 -> shared-memory operand stream, 991 MB per call)
  const int ntile = (g_gemm_wide_tiles && a.K % 256 == 0 && a.K >= 512 && nsplit <= 1) ? 256 : TM;
  const cuuint32_t bbox[3] = {KCH, (cuuint32_t)ntile, 1};
  if ((rc = gemm_tc_make_map(&mb, bt, 3, bdims, bstr, bbox))) return rc;
  GemmTcParams p = {};
  p.mode = 2; p.M = a.M; p.n_mtiles = (a.M + TM - 1) / TM; p.n_ntiles = (a.K + ntile - 1) / ntile; p.n_tile = ntile;
  const int ns = nsplit > 1 ? nsplit : 1;
  p.kchunks = (Np / KCH + ns - 1) / ns;                  // per split; partial (b, s) is matrix b * ns + s of c
  p.ksplit = ns;
  p.a_batched = 1; p.Kp = Np; p.N = a.K;
  p.y = a.c; p.y_stride_b = a.c_stride_b; p.ldy = a.ldc;
  if (a.ln_gamma != nullptr) {
    if (a.K != TM || a.B != 1 || ns != 1 || a.c_hi != nullptr || a.ln_beta == nullptr || a.ldc % 4 != 0 ||
        ((reinterpret_cast<uintptr_t>(a.ln_gamma) | reinterpret_cast<uintptr_t>(a.ln_beta) | reinterpret_cast<uintptr_t>(a.c_res)) & 15) != 0) {
      emip_set_error("gemm_nt_tc: the LayerNorm epilogue needs 128 output columns and 16-byte aligned gamma / beta / residual");
      return EMIP_ENOSYS;
    }
    p.ln_gamma = a.ln_gamma; p.ln_beta = a.ln_beta; p.ln_eps = a.ln_eps;
    p.res = a.c_res; p.res_stride_b = 0; p.ldr = a.ldc;
  }
  if (a.ln_hi != nullptr) {
    if (a.ln_gamma == nullptr || a.ln_lo == nullptr || a.ln_ld % 4 != 0 || ((reinterpret_cast<uintptr_t>(a.ln_hi) | reinterpret_cast<uintptr_t>(a.ln_lo)) & 7) != 0) {
      emip_set_error("gemm_nt_tc: ln_hi needs the LayerNorm epilogue and 8-byte aligned rows"); return EMIP_EINVAL;
    }
    p.ln_hi = static_cast<__nv_bfloat16*>(a.ln_hi); p.ln_lo = static_cast<__nv_bfloat16*>(a.ln_lo); p.ln_ld = a.ln_ld;
  }
  if (a.split_dst[0] != nullptr) {
    if (a.B != 1 || ns != 1 || ntile != TM || a.K > 4 * TM || a.K % TM != 0 || a.rowmap == nullptr || a.npix <= 0 || a.n_img <= 0 ||
        a.M != a.npix * a.n_img || a.ln_gamma != nullptr || a.c_hi != nullptr) {
      emip_set_error("gemm_nt_tc: bad window-ordered split output"); return EMIP_EINVAL;
    }
    for (int i = 0; i < 4; ++i) p.split_dst[i] = static_cast<__nv_bfloat16*>(a.split_dst[i]);
    p.rowmap = static_cast<const int2*>(a.rowmap); p.npix = a.npix; p.n_img = a.n_img; p.img_shift = a.img_shift;
    p.out_split = 3; p.ldg = 256;
    p.g_hi = p.split_dst[0]; p.g_lo = p.split_dst[0] + 128;
  }
  if (a.c_bias != nullptr) {
    if (a.ln_gamma != nullptr || a.split_dst[0] != nullptr || ns != 1 || (a.c_hi != nullptr && (reinterpret_cast<uintptr_t>(a.c_bias) & 15) != 0)) {
      emip_set_error("gemm_nt_tc: c_bias needs the plain fp32 or the bf16 hi | lo epilogue (bias 16-byte aligned there)");
      return EMIP_ENOSYS;
    }
    p.bias_n = a.c_bias;
  }
  if (a.c_hi != nullptr) {
    p.g_hi = static_cast<__nv_bfloat16*>(a.c_hi); p.g_lo = static_cast<__nv_bfloat16*>(a.c_lo);
    p.out_split = a.c_act ? 2 : 1; p.ldg = a.ldc_split;
  }
  return gemm_tc_launch(ma_hi, ma_lo, mb, p, a.B * ns, st);
}

// =====================================================================================================================
// f1 backward: the gradients of  out[b,o,p] = bias[o] + sum_{t,c} G[b][(o,t)][c] f0[b,c,p + d_t],
//                                G[b][(o,t)][c] = s sum_j Wp[(o,t)][j] f1[b,c,j],  s = 1/sqrt(C)      (conv_corr.cu)
// as five mode-2 launches of the tensor-core GEMM on K-major bf16 hi|lo operands (X9[b][(t,c)][p] = f0[b,c,p + d_t], the
// nine zero-padded shifts of f0, exists only as the split operand of GEMM (i)):
//   G / s      = Wp f1^T                   M = 9 O, N = 128,     K = HW      (recomputed: the forward keeps it in bf16 only)
//   (i)   dG   = dout X9^T                 M = O,   N = 9 * 128, K = HW      per sample
//   (ii)  dX9  = G^T dout                  M = 9 * 128, N = HW,  K = O       per sample; df0 = col2im(dX9)
//   (iii) dWp  = s sum_(b,c) dG f1         M = 9 O, N = HW,      K = 128 B   (batch folded into K)
//   (iv)  df1  = s sum_(o,t) dG^T Wp       M = 128 B, N = HW,    K = 9 O     (batch folded into M)
// r1 re-derived these with library ops (three fp32 SIMT sgemms + 16 cuDNN wgrad calls: 5.6 ms at B = 16).
namespace {

// dbias[o] = sum_{b,p} dout[b][o][p]
__global__ void __launch_bounds__(256)
f1_dbias_kernel(const float* __restrict__ dout, float* __restrict__ dbias, int B, int O, int P) {
  __shared__ float red[8];
  const int o = blockIdx.x, P4 = P >> 2;                 // P % 4 == 0 (checked by the caller)
  float s = 0.f;
  for (int i = threadIdx.x; i < B * P4; i += blockDim.x) {
    const int b = i / P4, q = i - b * P4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(dout + ((size_t)b * O + o) * P) + q);
    s += (v.x + v.y) + (v.z + v.w);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    dbias[o] = t;
  }
}
// Bt[b][(t,c)][hi | lo over p] = f0[b][c][p + d_t] (0 outside the image): the B operand of (i).
// grid (ceil(Pp / 2048), 9 * 128, B): 8 consecutive pixels per thread.
__global__ void __launch_bounds__(256)
f1_im2col_split_kernel(const float* __restrict__ f0, __nv_bfloat16* __restrict__ dst, int H, int W, int Pp) {
  const int b = blockIdx.z, tc_ = blockIdx.y, t = tc_ >> 7, c = tc_ & 127;
  const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (p0 >= Pp) return;
  const int dy = t / 3 - 1, dx = t % 3 - 1, P = H * W;
  const float* src = f0 + ((size_t)b * 128 + c) * P;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = p0 + i, y = p / W + dy, x = p % W + dx;
    v[i] = (p < P && y >= 0 && y < H && x >= 0 && x < W) ? __ldg(src + y * W + x) : 0.f;
  }
  __nv_bfloat16* d = dst + ((size_t)b * 9 * 128 + tc_) * 2 * Pp + p0;
  split8_store(v, d, d + Pp);
}
// df0[b][c][q] = s sum_t dX9[b][(t,c)][q - d_t]   (s: G is kept unscaled, the 1/sqrt(C) of (ii) is applied here)
__global__ void __launch_bounds__(256)
f1_col2im_kernel(const float* __restrict__ dx9, float* __restrict__ df0, int H, int W, float s_) {
  const int b = blockIdx.z, c = blockIdx.y, P = H * W;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P) return;
  const int qy = q / W, qx = q % W;
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int y = qy - (t / 3 - 1), x = qx - (t % 3 - 1);
    if (y >= 0 && y < H && x >= 0 && x < W) s += __ldg(dx9 + (((size_t)b * 9 + t) * 128 + c) * P + y * W + x);
  }
  df0[((size_t)b * 128 + c) * P + q] = s * s_;
}
// A of (iii): hi, lo [M1][Kp = 128 B] with A[m][b * 128 + c] = s dG[b][m][c]
__global__ void __launch_bounds__(128)
f1_gather_split_kernel(const float* __restrict__ dg, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int B, int M1,
                       float s) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 4), b = blockIdx.y;
  const int c0 = (threadIdx.x & 15) * 8;
  if (m >= M1) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(dg + ((size_t)b * M1 + m) * 128 + c0));
  const float4 c = __ldg(reinterpret_cast<const float4*>(dg + ((size_t)b * M1 + m) * 128 + c0 + 4));
  const float v[8] = {a.x * s, a.y * s, a.z * s, a.w * s, c.x * s, c.y * s, c.z * s, c.w * s};
  const size_t o = (size_t)m * (128 * B) + (size_t)b * 128 + c0;
  split8_store(v, hi + o, lo + o);
}
// Bt of (iv): [j][hi | lo over k = (o,t)] = s w[o][j][t], zero padded to Kp
__global__ void __launch_bounds__(256)
f1_weight_t_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int O, int N, int Kp, float s) {
  const int j = blockIdx.y;
  const int k0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (k0 >= Kp) return;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + i, o = k / 9, t = k - o * 9;
    v[i] = o < O ? __ldg(w + ((size_t)o * N + j) * 9 + t) * s : 0.f;
  }
  __nv_bfloat16* d = dst + (size_t)j * 2 * Kp + k0;
  split8_store(v, d, d + Kp);
}
// dweight[o][j][t] = dWp[(o,t)][j]
__global__ void __launch_bounds__(256)
f1_dweight_permute_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int O, int N) {
  const int o = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
    float* d = dw + ((size_t)o * N + j) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) d[t] = __ldg(dwp + ((size_t)o * 9 + t) * N + j);
  }
}

struct F1Bufs {
  float *g, *dg, *dx9, *dwp;
  __nv_bfloat16 *dsp, *fsp;   // dout and f1 split once as rows [.][row][hi | lo over the pixel axis] (pitch 2 Pp)
  char* ops;        // operand scratch, reused by every GEMM
  size_t ops_bytes;
};
size_t f1_ops_bytes(int B, int O, int P) {
  const size_t Pp = kpad(P), M1 = (size_t)O * 9, Ko = kpad(O), Kb = (size_t)128 * B, Km = kpad((int)M1);
  auto al = [](size_t x) { return emip_align_up(x, 1024); };
  size_t need = 0, t;
  t = al((size_t)B * 128 * 2 * Pp * 2);                                              // G: Bt = f1 rows
  need = t > need ? t : need;
  t = al((size_t)B * O * Pp * 2 * 2) + al((size_t)B * 9 * 128 * 2 * Pp * 2);         // (i): A = dout, Bt = X9
  need = t > need ? t : need;
  t = al((size_t)B * 9 * 128 * Ko * 2 * 2) + al((size_t)B * P * 2 * Ko * 2);         // (ii): A = G^T, Bt = dout token-major
  need = t > need ? t : need;
  t = al(M1 * Kb * 2 * 2) + al((size_t)P * 2 * Kb * 2);                              // (iii)
  need = t > need ? t : need;
  t = al(Kb * Km * 2 * 2) + al((size_t)P * 2 * Km * 2);                              // (iv)
  need = t > need ? t : need;
  return need;
}
size_t f1_carve(char* base, int B, int O, int P, F1Bufs* f) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += emip_align_up(bytes, 1024); return p; };
  const size_t M1 = (size_t)O * 9;
  F1Bufs t;
  t.g = reinterpret_cast<float*>(take((size_t)B * M1 * 128 * 4));
  t.dg = reinterpret_cast<float*>(take((size_t)B * M1 * 128 * 4));
  t.dx9 = reinterpret_cast<float*>(take((size_t)B * 9 * 128 * P * 4));
  t.dwp = reinterpret_cast<float*>(take(M1 * P * 4));
  t.dsp = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * O * 2 * kpad(P) * 2));
  t.fsp = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * 128 * 2 * kpad(P) * 2));
  t.ops_bytes = f1_ops_bytes(B, O, P);
  t.ops = take(t.ops_bytes);
  if (f) *f = t;
  return off;
}

// y[b][m][n] = sum_k A(b or shared)[m][k] Bt[b][n][k] on already split operands
// b_mn: bt is [batch][K rows][2 * Np] (n contiguous) and is read as an MN-major operand; else [batch][N][2 * Kp]
int f1_gemm(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, size_t a_ld, int a_batched, const __nv_bfloat16* bt, int M, int N,
            int Kp, int batch, float* y, long long y_stride_b, int ldy, cudaStream_t st, int b_mn = 0, int Krows = 0, int Np = 0) {
  CUtensorMap ma_hi, ma_lo, mb;
  int rc;
  const cuuint64_t adims[3] = {(cuuint64_t)Kp, (cuuint64_t)M, (cuuint64_t)(a_batched ? batch : 1)};
  const cuuint64_t astr[2] = {(cuuint64_t)a_ld * 2, (cuuint64_t)M * a_ld * 2};
  const cuuint32_t abox[3] = {KCH, TM, 1};
  if ((rc = gemm_tc_make_map(&ma_hi, a_hi, 3, adims, astr, abox))) return rc;
  if ((rc = gemm_tc_make_map(&ma_lo, a_lo, 3, adims, astr, abox))) return rc;
  if (b_mn) {
    const cuuint64_t bdims[3] = {(cuuint64_t)2 * Np, (cuuint64_t)Krows, (cuuint64_t)batch};
    const cuuint64_t bstr[2] = {(cuuint64_t)2 * Np * 2, (cuuint64_t)Krows * 2 * Np * 2};
    const cuuint32_t bbox[3] = {64, 64, 1};
    if ((rc = gemm_tc_make_map(&mb, bt, 3, bdims, bstr, bbox))) return rc;
  } else {
    const cuuint64_t bdims[3] = {(cuuint64_t)2 * Kp, (cuuint64_t)N, (cuuint64_t)batch};
    const cuuint64_t bstr[2] = {(cuuint64_t)2 * Kp * 2, (cuuint64_t)N * 2 * Kp * 2};
    const cuuint32_t bbox[3] = {KCH, TM, 1};
    if ((rc = gemm_tc_make_map(&mb, bt, 3, bdims, bstr, bbox))) return rc;
  }
  GemmTcParams p = {};
  p.mode = 2; p.M = M; p.n_mtiles = (M + TM - 1) / TM; p.n_ntiles = (N + TM - 1) / TM; p.n_tile = TM;
  p.kchunks = Kp / KCH;
  p.a_batched = a_batched; p.Kp = Kp; p.N = N;
  p.b_mn = b_mn; p.Np = Np;
  p.y = y; p.y_stride_b = y_stride_b; p.ldy = ldy;
  return gemm_tc_launch(ma_hi, ma_lo, mb, p, batch, st);
}
}  // namespace

size_t conv_corr_bwd_scratch_bytes(int B, int O, int P) { return f1_carve(nullptr, B, O, P, nullptr); }

int conv_corr_bwd_tc(const float* f0, const float* f1, const float* weight, const void* w_prep, long long w_prep_ld,
                     const float* dout, float* df0, float* df1, float* dweight, float* dbias, void* scratch, size_t scratch_bytes,
                     int B, int H, int W, int O, cudaStream_t st) {
  const int P = H * W, M1 = O * 9, Pp = kpad(P), Ko = kpad(O), Kb = 128 * B, Km = kpad(M1);
  const float s = 1.0f / sqrtf(128.0f);
  if (scratch == nullptr || scratch_bytes < conv_corr_bwd_scratch_bytes(B, O, P) || reinterpret_cast<uintptr_t>(scratch) % 1024 != 0) {
    emip_set_error("conv_corr_bwd: scratch too small or not 1024-byte aligned");
    return EMIP_ENOMEM;
  }
  if (P % 4 != 0) { emip_set_error("conv_corr_bwd: H * W must be a multiple of 4"); return EMIP_ENOSYS; }
  F1Bufs f;
  f1_carve(static_cast<char*>(scratch), B, O, P, &f);
  auto al = [](size_t x) { return emip_align_up(x, 1024); };
  int rc;
  if (dbias != nullptr) {
    f1_dbias_kernel<<<O, 256, 0, st>>>(dout, dbias, B, O, P);
    EMIP_CHECK_LAUNCH("conv_corr_bwd (dbias)");
  }
  // ---- G = s Wp f1^T  (A = the prepared weight of the forward, K-major over j)
  {
    // f1 and dout are split ONCE, as rows over the pixel axis: f1 rows are the K-major B operand of G and the MN-major B
    // operand of (iii); dout rows are the K-major A operand of (i) and the MN-major B operand of (ii)
    __nv_bfloat16* bt = f.fsp;
    split_rows_kernel<<<dim3((Pp + 2047) / 2048, 128, B), 256, 0, st>>>(f1, (long long)128 * P, P, nullptr, nullptr, nullptr, nullptr, bt,
                                                                        128, P, Pp);
    EMIP_CHECK_LAUNCH("conv_corr_bwd (f1 rows)");
    split_rows_kernel<<<dim3((Pp + 2047) / 2048, O, B), 256, 0, st>>>(dout, (long long)O * P, P, nullptr, nullptr, nullptr, nullptr, f.dsp,
                                                                      O, P, Pp);
    EMIP_CHECK_LAUNCH("conv_corr_bwd (dout rows)");
    const __nv_bfloat16* w_hi = static_cast<const __nv_bfloat16*>(w_prep);
    const __nv_bfloat16* w_lo = w_hi + (size_t)M1 * w_prep_ld;
    // the prepared weight has P valid columns per row (pitch w_prep_ld): columns >= P of the last K chunk are zero-filled
    // by TMA only if the map says so -- describe exactly P columns
    CUtensorMap ma_hi, ma_lo, mb;
    const cuuint64_t adims[3] = {(cuuint64_t)P, (cuuint64_t)M1, 1}, astr[2] = {(cuuint64_t)w_prep_ld * 2, (cuuint64_t)M1 * w_prep_ld * 2};
    const cuuint32_t abox[3] = {KCH, TM, 1};
    if ((rc = gemm_tc_make_map(&ma_hi, w_hi, 3, adims, astr, abox))) return rc;
    if ((rc = gemm_tc_make_map(&ma_lo, w_lo, 3, adims, astr, abox))) return rc;
    const cuuint64_t bdims[3] = {(cuuint64_t)2 * Pp, 128, (cuuint64_t)B}, bstr[2] = {(cuuint64_t)2 * Pp * 2, (cuuint64_t)128 * 2 * Pp * 2};
    const cuuint32_t bbox[3] = {KCH, TM, 1};
    if ((rc = gemm_tc_make_map(&mb, bt, 3, bdims, bstr, bbox))) return rc;
    GemmTcParams p = {};
    p.mode = 2; p.M = M1; p.n_mtiles = (M1 + TM - 1) / TM; p.n_ntiles = 1; p.n_tile = TM;
    p.kchunks = Pp / KCH;
    p.a_batched = 0; p.Kp = Pp; p.N = 128;
    p.y = f.g; p.y_stride_b = (long long)M1 * 128; p.ldy = 128;
    if ((rc = gemm_tc_launch(ma_hi, ma_lo, mb, p, B, st))) return rc;      // G / s: the scale is applied by its consumer
  }
  // ---- (i) dG[b] = dout[b] X9[b]^T
  {
    __nv_bfloat16* bt = reinterpret_cast<__nv_bfloat16*>(f.ops);
    f1_im2col_split_kernel<<<dim3((Pp + 2047) / 2048, 9 * 128, B), 256, 0, st>>>(f0, bt, H, W, Pp);
    EMIP_CHECK_LAUNCH("conv_corr_bwd (im2col)");
    if ((rc = f1_gemm(f.dsp, f.dsp + Pp, 2 * (size_t)Pp, 1, bt, O, 9 * 128, Pp, B, f.dg, (long long)M1 * 128, 9 * 128, st))) return rc;
  }
  // ---- (ii) dX9[b] = G[b]^T dout[b];  df0 = col2im
  {
    __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(f.ops);
    __nv_bfloat16* a_lo = a_hi + (size_t)B * 9 * 128 * Ko;
    split_w_kernel<<<dim3(split_w_blocks(9 * 128, Ko), 1, B), 256, 0, st>>>(f.g, (long long)M1 * 128, 9 * 128, 1, a_hi, a_lo, 9 * 128, O, Ko);
    EMIP_CHECK_LAUNCH("conv_corr_bwd (G^T)");
    if ((rc = f1_gemm(a_hi, a_lo, Ko, 1, f.dsp, 9 * 128, P, Ko, B, f.dx9, (long long)9 * 128 * P, P, st, 1, O, Pp))) return rc;
    f1_col2im_kernel<<<dim3((P + 255) / 256, 128, B), 256, 0, st>>>(f.dx9, df0, H, W, s);
    EMIP_CHECK_LAUNCH("conv_corr_bwd (col2im)");
  }
  // ---- (iii) dWp = s sum_(b,c) dG f1
  {
    __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(f.ops);
    __nv_bfloat16* a_lo = a_hi + (size_t)M1 * Kb;
    f1_gather_split_kernel<<<dim3((M1 + 7) / 8, B), 128, 0, st>>>(f.dg, a_hi, a_lo, B, M1, s);
    EMIP_CHECK_LAUNCH("conv_corr_bwd (dG gather)");
    if ((rc = f1_gemm(a_hi, a_lo, Kb, 0, f.fsp, M1, P, Kb, 1, f.dwp, 0, P, st, 1, Kb, Pp))) return rc;
    f1_dweight_permute_kernel<<<dim3((P + 255) / 256, O), 256, 0, st>>>(f.dwp, dweight, O, P);
    EMIP_CHECK_LAUNCH("conv_corr_bwd (dweight)");
  }
  // ---- (iv) df1 = s sum_(o,t) dG^T Wp
  {
    __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(f.ops);
    __nv_bfloat16* a_lo = a_hi + (size_t)Kb * Km;
    __nv_bfloat16* bt = reinterpret_cast<__nv_bfloat16*>(f.ops + al((size_t)Kb * Km * 2 * 2));
    split_w_kernel<<<dim3(split_w_blocks(128, Km), 1, B), 256, 0, st>>>(f.dg, (long long)M1 * 128, 128, 1, a_hi, a_lo, 128, M1, Km);
    EMIP_CHECK_LAUNCH("conv_corr_bwd (dG^T)");
    f1_weight_t_split_kernel<<<dim3((Km + 2047) / 2048, P), 256, 0, st>>>(weight, bt, O, P, Km, s);
    EMIP_CHECK_LAUNCH("conv_corr_bwd (weight^T)");
    if ((rc = f1_gemm(a_hi, a_lo, Km, 0, bt, Kb, P, Km, 1, df1, 0, P, st))) return rc;
  }
  return EMIP_OK;
}
