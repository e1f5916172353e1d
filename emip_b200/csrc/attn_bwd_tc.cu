// Backward of  O = softmax(Q K^T / sqrt(C)) V  (C = d_v = 128) on the 5th-gen tensor cores, for a5 (EMIP_long memory
// read, model/EMIP_long/LTM.py:49-68) and f2 (attention core of the GMFlow FeatureTransformer,
// model/EMIP_short/motion/gmflow/transformer.py:8-16, :46-105).  What autograd derives for those lines:
//   P = e^{S - L},  dV = P^T dO,  dP = dO V^T,  dS = P o (dP - D),  D = rowsum(dO o O),  dQ = dS K / sqrt(C),  dK = dS^T Q / sqrt(C)
// One kernel, three launches (attn_bwd_tc.cuh): dQ with rows = queries; dK and dV with rows = keys (the score tile is
// recomputed transposed, K Q^T, so that every contraction runs over TMEM columns / smem rows, never over TMEM lanes).
//
// CTA = one 128-row tile, persistent over (problem, row tile, column split) items, 576 threads:
//   warps 0..15 tile math + epilogue (thread <-> TMEM lane <-> row; warp w: lane quarter w%4, 32-column part w/4)
//   warp 16     TMA producer: the row operand x (for S) once per item (resident, 64 KB), then per 128-column tile the
//               operands of dP (g, z) and the column operand y through a 10-stage ring of 16 KB chunks.  (A first
//               version kept g resident too and had 6 stages: the y tile occupies 4 of them from UMMA-1a to UMMA-2,
//               so nothing of the next tile could be prefetched and every tile exposed the L2 latency: 16k cycles per
//               tile.  Streaming g costs 64 KB more L2 traffic per tile and hides it under the math.)
//   warp 17     UMMA issuer
// TMEM: S [0,128) | dP [128,256) | W as bf16 hi [256,320) + lo [320,384) | accumulator [384,512).
//   UMMA-1b (SS): dP = g.hi z.hi^T + g.lo z.hi^T + g.hi z.lo^T          (not in PV mode)
//   UMMA-1a (SS): S  = x.hi y.hi^T + x.lo y.hi^T + x.hi y.lo^T
//   math:         W  = 2^{S c2 - L} (dP - D), split into bf16 hi + lo, back to TMEM
//   UMMA-2  (TS): acc += W.hi B.hi + W.lo B.hi + W.hi B.lo,  B = the y tile (ROW, COL: it stays in its ring stages from
//                 UMMA-1a to UMMA-2) or the z tile (PV), read as an MN-major operand straight from the token-major tile
// All operands are 3-term split bf16 with fp32 accumulation, as in the forward.  In ROW / COL mode the MMAs of a tile
// run strictly in sequence with its math (UMMA-1, math, UMMA-2; S, dP and W are single-buffered in TMEM and the y tile
// occupies 4 of the 10 ring stages until UMMA-2, so the 12 chunks of the next tile cannot all be staged), only the
// operand loads overlap: the tensor pipe idles during the math.  PV mode issues S(t + 1) ahead of
// UMMA-2(t) like the forward kernel.
#include "common.cuh"
#include "pair_common.cuh"
#include "tc_common.cuh"
#include "attn_bwd_tc.cuh"
#include "attn_tc.cuh"

namespace {
using namespace tc;

constexpr int TN = 128;
constexpr int CH_ELEMS = 64;
constexpr int CHUNK_BYTES = TM * 128;   // 16 KB
constexpr int STAGES = 10;
constexpr int NMATH = 16;
constexpr int NTHREADS = (NMATH + 2) * 32;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_S = 0, COL_DP = 128, COL_W = 256, COL_ACC = 384;

constexpr int OFF_X = 0;                                  // 4 chunks: hi[0:64] hi[64:128] lo[0:64] lo[64:128]
constexpr int OFF_RING = OFF_X + 4 * CHUNK_BYTES;
constexpr int OFF_TAB = OFF_RING + STAGES * CHUNK_BYTES;  // [2][TN] floats: L log2(e), D of the current column tile
constexpr int OFF_BAR = OFF_TAB + 2 * TN * 4;
constexpr int NBAR = 2 + 2 * STAGES + 6;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
static_assert(STAGES % 2 == 0, "MN-major operands need their two 64-channel chunks in adjacent ring stages");

struct BParams {
  AttnBwdTcArgs a;
  float inv_sqrt_c;
  unsigned long long* prof;   // optional [gridDim.x][16] cycle counters (emip_attn_tc_set_profile_buffer)
};

template <bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                   const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_z,
                   const __grid_constant__ BParams bp) {
  const AttnBwdTcArgs& p = bp.a;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + OFF_BAR;
  const uint32_t x_full = bar0, x_empty = bar0 + 8;
  auto r_full = [&](int s) { return bar0 + 16 + 8 * s; };
  auto r_empty = [&](int s) { return bar0 + 16 + 8 * (STAGES + s); };
  const uint32_t s_full = bar0 + 16 + 8 * (2 * STAGES), s_empty = s_full + 8, w_full = s_full + 16, w_empty = s_full + 24,
                 acc_full = s_full + 32, acc_empty = s_full + 40;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_dp = p.mode != ATTN_BWD_PV;
  const int nrt = (p.nr + TM - 1) / TM;
  const int nkt = (p.nc + TN - 1) / TN;
  const int ns = p.ksplit > 1 ? p.ksplit : 1;
  const int n_items = p.nb * nrt * ns;

  if (threadIdx.x == 0) {
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(r_full(s), 1); mbar_init(r_empty(s), 1); }
    mbar_init(s_full, 1);
    mbar_init(s_empty, NMATH * 32);
    mbar_init(w_full, NMATH * 32);
    mbar_init(w_empty, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, NMATH * 32);
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
      mbar_wait(x_empty, (it & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(x_full, 4 * CHUNK_BYTES);
        for (int c = 0; c < 4; ++c) tma_load_3d(sbase + OFF_X + c * CHUNK_BYTES, &map_x, x_full, c * CH_ELEMS, rt * TM, prob);
      }
      for (int t = kb; t < ke; ++t) {
        // ring order = consumption order of the issuer: (g, z for dP; y for S and UMMA-2) or, in PV mode, (y for S; z for UMMA-2)
        if (has_dp) {
          for (int c = 0; c < 4; ++c) push(&map_g, c * CH_ELEMS, rt * TM, prob);
          for (int c = 0; c < 4; ++c) push(&map_z, c * CH_ELEMS, t * TN, prob);
        }
        if (has_dp || t == kb)
          for (int c = 0; c < 4; ++c) push(&map_y, c * CH_ELEMS, t * TN, prob);
        if (!has_dp) {                    // PV is pipelined like the forward: y(t + 1) is consumed before z(t)
          if (t + 1 < ke)
            for (int c = 0; c < 4; ++c) push(&map_y, c * CH_ELEMS, (t + 1) * TN, prob);
          for (int c = 0; c < 4; ++c) push(&map_z, c * CH_ELEMS, t * TN, prob);
        }
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
    const uint32_t idesc_mn = idesc_full | (1u << 16);    // b_major = MN
    const uint32_t t_w = tmem_base + COL_W, t_acc = tmem_base + COL_ACC;
    long long w_rf = 0, w_se = 0, w_wf = 0, w_ae = 0, w_xf = 0;
    const long long t_begin = PROF ? clock64() : 0;
    // D[128 x n] = A.hi B.hi^T + A.lo B.hi^T + A.hi B.lo^T with A resident at a_off and B in the next four ring stages;
    // `release` hands the stages back to the producer (the y tile of ROW / COL mode is kept for UMMA-2)
    auto mma1 = [&](int a_off, uint32_t d_col, uint32_t idesc, bool release) {
      const uint64_t ad = make_kmajor_sw128_desc(sbase + a_off);
      for (int c = 0; c < 4; ++c) {
        prof_add<PROF>(w_rf, mbar_wait(r_full(stage), phase));
        tc_fence_after();
        if (leader) {
          const uint64_t kd = make_kmajor_sw128_desc(sbase + OFF_RING + stage * CHUNK_BYTES);
          const uint64_t a_hi = ad + (uint64_t)((c & 1) * (CHUNK_BYTES >> 4));
          const uint64_t a_lo = a_hi + (uint64_t)(2 * (CHUNK_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + d_col, a_hi + 2 * k, kd + 2 * k, idesc, (c | k) ? 1u : 0u);
          if (c < 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + d_col, a_lo + 2 * k, kd + 2 * k, idesc, 1u);
          }
          if (release) umma_commit(r_empty(stage));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    // dP = g.hi z.hi^T + g.lo z.hi^T + g.hi z.lo^T with both operands in the next eight ring stages (g: 4, z: 4)
    auto mma_dp = [&](uint32_t idesc) {
      int sg[4], sz[4];
      for (int c = 0; c < 8; ++c) {
        prof_add<PROF>(w_rf, mbar_wait(r_full(stage), phase));
        if (c < 4) sg[c] = stage; else sz[c - 4] = stage;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      tc_fence_after();
      if (leader) {
        uint64_t gd[4], zd[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          gd[c] = make_kmajor_sw128_desc(sbase + OFF_RING + sg[c] * CHUNK_BYTES);
          zd[c] = make_kmajor_sw128_desc(sbase + OFF_RING + sz[c] * CHUNK_BYTES);
        }
        const uint32_t d = tmem_base + COL_DP;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                     // channel halves [0:64], [64:128]
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, gd[h] + 2 * k, zd[h] + 2 * k, idesc, (h | k) ? 1u : 0u);          // g.hi z.hi
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, gd[2 + h] + 2 * k, zd[h] + 2 * k, idesc, 1u);                    // g.lo z.hi
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, gd[h] + 2 * k, zd[2 + h] + 2 * k, idesc, 1u);                    // g.hi z.lo
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) { umma_commit(r_empty(sg[c])); umma_commit(r_empty(sz[c])); }
      }
      __syncwarp();
    };
    // acc += W B with B = the token-major tile in ring stages s0 .. s0 + 3 (mod STAGES; pairs stay adjacent)
    auto mma2 = [&](int s0, bool first, bool last) {
      for (int half = 0; half < 2; ++half) {
        const int sa = (s0 + 2 * half) % STAGES;
        if (leader) {
          const uint64_t bd = make_mnmajor_sw128_desc(sbase + OFF_RING + sa * CHUNK_BYTES, CHUNK_BYTES);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t b = bd + (uint64_t)(kk * (2048 >> 4));
            umma_bf16_ts(t_acc, t_w + (uint32_t)(8 * kk), b, idesc_mn, (first && half == 0 && kk == 0) ? 0u : 1u);   // W.hi B.(hi|lo)
            if (half == 0) umma_bf16_ts(t_acc, t_w + 64 + (uint32_t)(8 * kk), b, idesc_mn, 1u);                     // W.lo B.hi
          }
          umma_commit(r_empty(sa));
          umma_commit(r_empty(sa + 1));
          if (half == 1) {
            umma_commit(w_empty);
            if (last) umma_commit(acc_full);
          }
        }
        __syncwarp();
      }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int ks = item % ns;
      const int kb = ks * nkt / ns, ke = (ks + 1) * nkt / ns;
      prof_add<PROF>(w_xf, mbar_wait(x_full, it & 1));
      tc_fence_after();
      if (!has_dp) {
        // PV (dV) has no dP: S(t + 1) is issued ahead of UMMA-2(t), as in the forward kernel, so the tensor pipe works
        // on the next score tile while the math warps turn the current one into W.  (Measured gain is small, 8970 ->
        // 8460 cycles per tile: the math phase itself, ~3700 cycles per tile next to a running UMMA, bounds this mode.)
        auto issue_s = [&](int t, uint32_t T) {
          prof_add<PROF>(w_se, mbar_wait(s_empty, (T & 1) ^ 1));
          tc_fence_after();
          mma1(OFF_X, COL_S, (t == nkt - 1) ? idesc_tail : idesc_full, true);
          if (leader) {
            umma_commit(s_full);
            if (t == ke - 1) umma_commit(x_empty);
          }
          __syncwarp();
        };
        issue_s(kb, tile);
        for (int t = kb; t < ke; ++t, ++tile) {
          if (t + 1 < ke) issue_s(t + 1, tile + 1);
          const int bs = stage;
          for (int c = 0; c < 4; ++c) {
            prof_add<PROF>(w_rf, mbar_wait(r_full(stage), phase));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          prof_add<PROF>(w_wf, mbar_wait(w_full, tile & 1));
          if (t == kb) prof_add<PROF>(w_ae, mbar_wait(acc_empty, (it & 1) ^ 1));
          tc_fence_after();
          mma2(bs, t == kb, t == ke - 1);
        }
        continue;
      }
      for (int t = kb; t < ke; ++t, ++tile) {
        const uint32_t idesc = (t == nkt - 1) ? idesc_tail : idesc_full;
        prof_add<PROF>(w_se, mbar_wait(s_empty, (tile & 1) ^ 1));       // the math warps hold the previous S / dP tiles in registers
        tc_fence_after();
        mma_dp(idesc);
        const int ys = stage;                             // first ring stage of the y tile
        mma1(OFF_X, COL_S, idesc, false);
        if (leader) {
          umma_commit(s_full);
          if (t == ke - 1) umma_commit(x_empty);          // x and g are read by UMMA-1 only
        }
        __syncwarp();
        const int bs = ys;
        prof_add<PROF>(w_wf, mbar_wait(w_full, tile & 1));
        if (t == kb) prof_add<PROF>(w_ae, mbar_wait(acc_empty, (it & 1) ^ 1));  // the epilogue of the previous item has drained the accumulator
        tc_fence_after();
        mma2(bs, t == kb, t == ke - 1);
      }
    }
    if (PROF && bp.prof && leader) {
      unsigned long long* o = bp.prof + blockIdx.x * 16;
      o[8] = w_se; o[9] = w_rf; o[10] = w_wf; o[11] = w_ae; o[12] = w_xf; o[13] = clock64() - t_begin;
    }
    __syncwarp();
  } else {
    // ===================== tile math + epilogue =====================
    const int quarter = warp & 3, part = warp >> 2;
    float* tab = reinterpret_cast<float*>(smem + OFF_TAB);
    const uint32_t tab_u32 = sbase + OFF_TAB;
    const float LOG2E = 1.4426950408889634f;
    const float c2 = LOG2E * bp.inv_sqrt_c;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int cb = part * 32;
    const bool row_stats = p.mode == ATTN_BWD_ROW;
    uint32_t tile = 0, it = 0;
    long long w_sf = 0, t_tab = 0, t_ld = 0, t_m = 0, w_we = 0, t_st = 0, w_af = 0;
    const long long t_begin = PROF ? clock64() : 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int ks = item % ns, rt = (item / ns) % nrt, prob = item / (ns * nrt);
      const int kb = ks * nkt / ns, ke = (ks + 1) * nkt / ns;
      const int row = rt * TM + quarter * 32 + lane;
      const bool row_ok = row < p.nr;
      float rL = 0.f, rD = 0.f;
      if (row_stats && row_ok) {
        const size_t si = p.win.enabled ? p.win.pixel(prob, row) : (size_t)prob * p.nr + row;
        rL = __ldg(p.lse + si) * LOG2E;
        rD = __ldg(p.dsum + si);
      }
      // per-column L log2(e) / D of key tile kt, one value per thread (threads 0..255); fetched one tile ahead so that
      // the global-load latency is not paid between the two barriers below
      auto col_stat = [&](int kt) -> float {
        float v = 0.f;
        if (!row_stats && threadIdx.x < 2 * TN && kt < ke) {
          const int k = threadIdx.x / TN, col = kt * TN + threadIdx.x % TN;
          if (col < p.nc) {
            const size_t si = p.win.enabled ? p.win.pixel(prob, col) : (size_t)prob * p.nc + col;
            if (k == 0) v = __ldg(p.lse + si) * LOG2E;
            else if (has_dp) v = __ldg(p.dsum + si);
          }
        }
        return v;
      };
      float stat_next = col_stat(kb);
      for (int kt = kb; kt < ke; ++kt, ++tile) {
        const int col_base = kt * TN;
        const long long cA = PROF ? clock64() : 0;
        if (!row_stats) {
          // this tile's table -> smem (single buffer: every warp has finished the previous tile)
          asm volatile("bar.sync 1, %0;" ::"n"(NMATH * 32) : "memory");
          if (threadIdx.x < 2 * TN) tab[threadIdx.x] = stat_next;
          asm volatile("bar.sync 1, %0;" ::"n"(NMATH * 32) : "memory");
          stat_next = col_stat(kt + 1);
        }
        if (PROF) t_tab += clock64() - cA;
        prof_add<PROF>(w_sf, mbar_wait(s_full, tile & 1));
        long long c0 = PROF ? clock64() : 0;
        tc_fence_after();
        const int nv = p.nc - (col_base + cb);              // valid columns of my part (warp-uniform)
        // two halves of 16 columns: 32 + 16 live registers instead of 64 + 32 (the first version spilled 250 bytes per
        // thread into an L1 that the 227 KB shared-memory carve-out leaves almost empty)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t rs[16], rp[16];
          tmem_ld16_async(tmem_base + lane_base + COL_S + (uint32_t)(cb + 16 * hf), rs);
          if (has_dp) tmem_ld16_async(tmem_base + lane_base + COL_DP + (uint32_t)(cb + 16 * hf), rp);
          tmem_wait16(rs);
          if (has_dp) tmem_wait16(rp);
          if (hf == 1) {
            tc_fence_before();
            mbar_arrive(s_empty);                           // S and dP are in registers: the next UMMA-1 may overwrite them
          }
          uint32_t whi[8], wlo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int cq = 16 * hf + 4 * q;                 // first column (within my part) of this group of four
            float w4[4];
            if (row_stats) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float pe = ex2f(fmaf(__uint_as_float(rs[4 * q + e]), c2, -rL));
                w4[e] = pe * (__uint_as_float(rp[4 * q + e]) - rD);
              }
            } else {
              const float4 l4 = lds128(tab_u32 + (uint32_t)((cb + cq) * 4));
              const float la[4] = {l4.x, l4.y, l4.z, l4.w};
              if (has_dp) {
                const float4 d4 = lds128(tab_u32 + (uint32_t)((TN + cb + cq) * 4));
                const float da[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  w4[e] = ex2f(fmaf(__uint_as_float(rs[4 * q + e]), c2, -la[e])) * (__uint_as_float(rp[4 * q + e]) - da[e]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) w4[e] = ex2f(fmaf(__uint_as_float(rs[4 * q + e]), c2, -la[e]));
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (!(row_ok && cq + e < nv)) w4[e] = 0.f;
            split_bf16x2_alu(w4[0], w4[1], whi[2 * q], wlo[2 * q]);
            split_bf16x2_alu(w4[2], w4[3], whi[2 * q + 1], wlo[2 * q + 1]);
          }
          if (hf == 0) {
            prof_add<PROF>(w_we, mbar_wait(w_empty, (tile & 1) ^ 1));     // UMMA-2 of the previous tile has retired: W may be overwritten
            tc_fence_after();
          }
          const uint32_t t_whi = tmem_base + lane_base + COL_W + (uint32_t)(part * 16 + 8 * hf);
          tmem_st8(t_whi, whi);
          tmem_st8(t_whi + 64, wlo);
        }
        if (PROF) t_m += clock64() - c0;
        if (PROF) c0 = clock64();
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(w_full);
        if (PROF) t_st += clock64() - c0;
      }
      // ---- epilogue: accumulator -> global
      prof_add<PROF>(w_af, mbar_wait(acc_full, it & 1));
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32_async(tmem_base + lane_base + COL_ACC + (uint32_t)cb, r);
        tmem_wait(r);
        tc_fence_before();
        mbar_arrive(acc_empty);
        const float sc = has_dp ? bp.inv_sqrt_c : 1.0f;
        float* OUT = ns == 1 ? p.out + (size_t)prob * p.out_stride_b : p.part + ((size_t)ks * p.nb + prob) * p.nr * 128;
        if (row_ok) {
          if (p.out_layout == EMIP_LAYOUT_NC) {
            size_t orow = (size_t)row;
            if (p.win.enabled) {                          // scatter: token `row` of block (prob / B) of image (prob % B)
              OUT = p.out;
              orow = p.win.pixel(prob, row);
            }
            float4* dst = reinterpret_cast<float4*>(OUT + orow * 128 + cb);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              dst[q] = make_float4(__uint_as_float(r[4 * q]) * sc, __uint_as_float(r[4 * q + 1]) * sc,
                                   __uint_as_float(r[4 * q + 2]) * sc, __uint_as_float(r[4 * q + 3]) * sc);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) OUT[(size_t)(cb + i) * p.nr + row] = __uint_as_float(r[i]) * sc;
          }
        }
      }
    }
    if (PROF && bp.prof && threadIdx.x == 0) {
      unsigned long long* o = bp.prof + blockIdx.x * 16;
      o[0] = w_sf; o[1] = t_tab; o[2] = t_ld; o[3] = t_m; o[4] = w_we; o[5] = t_st; o[6] = w_af; o[7] = clock64() - t_begin;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NMATH + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// out[b][i] = sum_s part[s][b][i]
__global__ void __launch_bounds__(256)
sum_parts_kernel(const float* __restrict__ part, float* __restrict__ out, long long out_stride_b, int ns, int nb, size_t per) {
  const int b = blockIdx.y;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < ns; ++s) acc += __ldcs(part + ((size_t)s * nb + b) * per + i);
    out[(size_t)b * out_stride_b + i] = acc;
  }
}

// D[b][r] = sum_c dO[b][r][c] O[b][r][c]
__global__ void __launch_bounds__(256)
dsum_nc_kernel(const float* __restrict__ d_o, long long do_stride_b, const float* __restrict__ o, long long o_stride_b,
               float* __restrict__ dsum, int n) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= n) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(d_o + (size_t)b * do_stride_b + (size_t)r * 128) + lane);
  const float4 c = __ldg(reinterpret_cast<const float4*>(o + (size_t)b * o_stride_b + (size_t)r * 128) + lane);
  const float s = warp_sum(a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w);
  if (lane == 0) dsum[(size_t)b * n + r] = s;
}
__global__ void __launch_bounds__(128)
dsum_cn_kernel(const float* __restrict__ d_o, long long do_stride_b, const float* __restrict__ o, long long o_stride_b,
               float* __restrict__ dsum, int n) {
  const int b = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float* a = d_o + (size_t)b * do_stride_b + r;
  const float* c = o + (size_t)b * o_stride_b + r;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int ch = 0; ch < 128; ++ch) s[ch & 3] = fmaf(__ldg(a + (size_t)ch * n), __ldg(c + (size_t)ch * n), s[ch & 3]);
  dsum[(size_t)b * n + r] = (s[0] + s[1]) + (s[2] + s[3]);
}

}  // namespace

int attn_bwd_tc(const AttnBwdTcArgs& a, cudaStream_t st) {
  if (a.nb == 0 || a.nr == 0) return EMIP_OK;
  if (a.nc < 1 || a.mode < 0 || a.mode > 2) { emip_set_error("attn_bwd_tc: bad arguments"); return EMIP_EINVAL; }
  const int nkt = (a.nc + TN - 1) / TN;
  if (a.ksplit > nkt) { emip_set_error("attn_bwd_tc: ksplit %d exceeds the %d column tiles", a.ksplit, nkt); return EMIP_EINVAL; }
  if (a.win.enabled && (a.ksplit > 1 || a.out_layout != EMIP_LAYOUT_NC)) { emip_set_error("attn_bwd_tc: the window scatter needs ksplit == 1 and the NC layout"); return EMIP_EINVAL; }
  if (a.ksplit > 1 && a.part == nullptr) { emip_set_error("attn_bwd_tc: ksplit needs the partial buffer"); return EMIP_EINVAL; }
  CUtensorMap mx, my, mg, mz;
  int rc;
  if ((rc = make_bf16_map(&mx, a.x_split, 256, (uint64_t)a.nr, (uint64_t)a.nb, 512, (uint64_t)a.nr * 512))) return rc;
  if ((rc = make_bf16_map(&my, a.y_split, 256, (uint64_t)a.nc, (uint64_t)a.nb, 512, (uint64_t)a.nc * 512))) return rc;
  if ((rc = make_bf16_map(&mz, a.z_split, 256, (uint64_t)a.nc, (uint64_t)a.nb, 512, (uint64_t)a.nc * 512))) return rc;
  mg = mx;
  if (a.mode != ATTN_BWD_PV &&
      (rc = make_bf16_map(&mg, a.g_split, 256, (uint64_t)a.nr, (uint64_t)a.nb, 512, (uint64_t)a.nr * 512)))
    return rc;
  if (int rc__ = emip_func_max_smem((const void*)(attn_bwd_tc_kernel<false>), SMEM_BYTES)) return rc__;
  if (int rc__ = emip_func_max_smem((const void*)(attn_bwd_tc_kernel<true>), SMEM_BYTES)) return rc__;
  BParams bp;
  bp.a = a;
  bp.inv_sqrt_c = 1.0f / a.sqrt_c;
  bp.prof = attn_tc_profile_buffer();
  const int nrt = (a.nr + TM - 1) / TM;
  long long grid = (long long)a.nb * nrt * (a.ksplit > 1 ? a.ksplit : 1);
  if (grid > emip_num_sms()) grid = emip_num_sms();
  if (bp.prof) attn_bwd_tc_kernel<true><<<(int)grid, NTHREADS, SMEM_BYTES, st>>>(mx, my, mg, mz, bp);   // diagnostics build
  else attn_bwd_tc_kernel<false><<<(int)grid, NTHREADS, SMEM_BYTES, st>>>(mx, my, mg, mz, bp);
  EMIP_CHECK_LAUNCH("attn_bwd_tc");
  return EMIP_OK;
}

int attn_bwd_tc_sum(const AttnBwdTcArgs& a, cudaStream_t st) {
  if (a.nb == 0 || a.nr == 0 || a.ksplit <= 1) return EMIP_OK;
  sum_parts_kernel<<<dim3(128, a.nb), 256, 0, st>>>(a.part, a.out, a.out_stride_b, a.ksplit, a.nb, (size_t)a.nr * 128);
  EMIP_CHECK_LAUNCH("attn_bwd_tc_sum");
  return EMIP_OK;
}

int attn_dsum(const float* d_o, long long do_stride_b, const float* o, long long o_stride_b, float* dsum, int nb, int n, int layout,
              cudaStream_t st) {
  if (nb == 0 || n == 0) return EMIP_OK;
  if (layout == EMIP_LAYOUT_NC) dsum_nc_kernel<<<dim3((n + 7) / 8, nb), 256, 0, st>>>(d_o, do_stride_b, o, o_stride_b, dsum, n);
  else dsum_cn_kernel<<<dim3((n + 127) / 128, nb), 128, 0, st>>>(d_o, do_stride_b, o, o_stride_b, dsum, n);
  EMIP_CHECK_LAUNCH("attn_dsum");
  return EMIP_OK;
}
