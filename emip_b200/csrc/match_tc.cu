// K1 / K2 forward on the 5th-gen tensor cores (sm_100a): fused
//   S = X Y^T / sqrt(C)  ->  online row softmax  ->  expectation of a 2-vector per column
// for a1 global matching (matching.py:16-39, both directions in one launch, the
// backward direction being the same problem with the operand roles swapped) and
// a2 flow-propagation attention (transformer.py:528-531).  The N x N score matrix
// lives only in TMEM; it reaches HBM only when the caller asks for `corr`.
//
// Numerics: fp32 features are split once into bf16 hi + bf16 lo
// (match_tc_split) and S is accumulated in fp32 as  hi.hi + lo.hi + hi.lo
// (24 UMMA K16-steps per 128x128 tile, one TMEM accumulator): S rel. error ~4e-6,
// flow rel-L2 ~1e-5 versus fp32 SGEMM.  terms=1 keeps only hi.hi (bf16 inference
// mode, 2e-2 tolerance).  Softmax statistics are always fp32.
//
// CTA = 576 threads, persistent, one CTA per SM (225 KB smem, all 512 TMEM columns).
// Scheduling is stream-K: the units (item, key tile) are dealt to the CTAs in equal contiguous spans, so every SM gets
// the same amount of MMA work whatever the batch (r1s: item-granular scheduling left 13 % of the SM-time idle at c2,
// 256 items on 148 SMs).  A work item is (problem, PAIR of 128-row query tiles): every key chunk that TMA
// brings in is used by both query tiles, which halves the L2->SM operand traffic
// per MMA (1 MB per 256 rows; the first version of this kernel, one tile per CTA, asked
// L2 for 11 TB/s at the MMA-bound rate).
//   warps 0..7   softmax, query tile 0: warp w <-> TMEM lane quarter w%4, column half (w/4)%2 of every key tile
//   warps 8..15  softmax, query tile 1      (thread <-> TMEM lane <-> one row; no shuffles; the two column halves
//                of a row keep independent online-softmax states that are merged once per work item)
//   warp 16      TMA producer   (2 x 64 KB query tiles per item; key chunks through a 4-stage ring)
//   warp 17      UMMA issuer    (one elected lane; owns the TMEM allocation)
// 16 softmax warps = 4 per SM sub-partition: with 8 (one query tile per warp group, r1d) each warp ran at IPC 0.3
// and a tile took 3500 cycles against 3072 of UMMA time per key step.
// The two control warps have the HIGHEST warp ids on purpose: the SM sub-partition arbiter favours
// high warp ids, and an issuer that shares its sub-partition with two busy softmax warps at low
// priority could only issue one UMMA per ~128 cycles (ncu r1c: tensor pipe 42 % active, softmax
// warps stalled on s_full).
// Each group has two 128-column accumulators in TMEM, so the MMAs of key tile t+1
// overlap the softmax of tile t.  `corr` is emitted by the softmax warps through
// a swizzled smem stage and TMA bulk stores (no strided st.global).
#include "common.cuh"
#include "match_tc.cuh"
#include "pair_common.cuh"
#include "tc_common.cuh"
#include <cuda.h>

namespace {
using namespace tc;

constexpr int TN = 128;                 // key columns per tile (UMMA N)
constexpr int CH_ELEMS = 64;            // bf16 per 128-byte swizzle row
constexpr int CHUNK_BYTES = TM * 128;   // one [128 rows x 128 B] SW128 box = 16 KB
constexpr int NCHUNK = 4;               // hi[0:64] hi[64:128] lo[0:64] lo[64:128]
constexpr int MAXK = 2048;              // value table capacity (columns)
constexpr uint32_t TMEM_COLS = 512;
constexpr int R_BYTES = 32768;          // region R, see below

// Shared-memory layout.  (A TS-mode variant with the query tiles in TMEM was tried in r1m: tcgen05.cp.128x256b and
// the TS-mode UMMA behave as expected -- tools/probe/ -- but the four 128-column accumulators already fill the 512
// TMEM columns, so it would cost the accumulator double-buffering.)
// Region R (32 KB) is the score-emission stage (16 x 2 KB) in matching mode, and the value table (16 KB) plus the
// end-of-item merge slots in attention mode -- the two uses never coexist (attention emits no scores, matching
// generates the pixel grid analytically and needs no table).
constexpr int STAGES = 4;
constexpr int OFF_Q = 0;                                        // [2 tiles][4 chunks][16 KB]
constexpr int OFF_K = OFF_Q + 2 * NCHUNK * CHUNK_BYTES;         // ring
constexpr int OFF_R = OFF_K + STAGES * CHUNK_BYTES;
constexpr int OFF_STG = OFF_R;
constexpr int OFF_V = OFF_R;
constexpr int OFF_MERGE = OFF_R + 2 * MAXK * 4;                 // attention mode: 16 x 512 B
constexpr int OFF_BAR = OFF_R + R_BYTES;
constexpr int NBAR = 3 + 2 * STAGES + 8;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;                // + slack to align the base to 1024 B
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
static_assert(2 * MAXK * 4 + 16 * 512 <= R_BYTES, "region R too small");

struct KParams {
  const float* v;
  long long v_stride_b;
  float* out;
  float* lse;
  float* s_out;          // direct-store fallback for the scores (used when the TMA store map cannot be built)
  int nb, nq, nk, y_shift, y_mod, s_first, s_count;
  int grid_w, sub_grid, terms, s_mode;   // s_mode: 0 none, 1 TMA bulk store, 2 direct st.global
  float inv_sqrt_c;
  unsigned long long* prof;   // optional [gridDim.x][8] wait-cycle counters (emip_match_tc_set_profile_buffer)
  // stream-K scheduling: the (item, key tile) units are dealt to the CTAs in equal contiguous spans; a row whose keys
  // are covered by several CTAs leaves its online-softmax state in `part` and the last CTA to arrive merges them
  float4* part;          // [n_items][pmax][2 tiles][128 rows] (m, l, sx, sy)
  unsigned* cnt;         // [n_items] arrival counters, zero before the launch and left at zero
  int pmax;
};

// Online-softmax update of one row with 32 score columns (raw accumulator values, scale folded into c2).
template <bool FULL>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&r)[32], int nv, uint32_t vx, uint32_t vy, float c2,
                                              float& m, float& l, float& sx, float& sy) {
  float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float s = __uint_as_float(r[i]);
    cm[i & 3] = (FULL || i < nv) ? fmaxf(cm[i & 3], s) : cm[i & 3];
  }
  const float m_new = fmaxf(fmaxf(m, fmaxf(cm[0], cm[1])), fmaxf(cm[2], cm[3]));
  const float corr = ex2f((m - m_new) * c2);         // first chunk: ex2(-inf) = 0
  const float mb = m_new * c2;
  float ls[4] = {0.f, 0.f, 0.f, 0.f}, lx[4] = {0.f, 0.f, 0.f, 0.f}, ly[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i4 = 0; i4 < 8; ++i4) {
    const float4 ax = lds128(vx + 16 * i4), ay = lds128(vy + 16 * i4);   // value table in smem (warp-wide broadcast)
    const float xs[4] = {ax.x, ax.y, ax.z, ax.w}, ys[4] = {ay.x, ay.y, ay.z, ay.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float e = ex2f(fmaf(__uint_as_float(r[i4 * 4 + k]), c2, -mb));
      if (!FULL) e = (i4 * 4 + k < nv) ? e : 0.f;
      ls[k] += e;
      lx[k] = fmaf(e, xs[k], lx[k]);
      ly[k] = fmaf(e, ys[k], ly[k]);
    }
  }
  l = fmaf(l, corr, (ls[0] + ls[1]) + (ls[2] + ls[3]));
  sx = fmaf(sx, corr, (lx[0] + lx[1]) + (lx[2] + lx[3]));
  sy = fmaf(sy, corr, (ly[0] + ly[1]) + (ly[2] + ly[3]));
  m = m_new;
}

// Same update for the analytic pixel grid (geometry.py:5-21) when grid_w >= 32 and grid_w % 4 == 0: the 32 columns
// start at grid position (x0, y0) and wrap to the next grid row at most once, at a multiple of four (brk4 * 4):
//   x_i = x0 + i - W [i >= brk],  y_i = y0 + [i >= brk]
//   sum e_i x_i = x0 L + sum i e_i - W H,   sum e_i y_i = y0 L + H,   L = sum e_i,  H = sum_{i >= brk} e_i
// so the inner loop is one FFMA (scale), one MUFU, one FADD (groups of four -> H comes from 8 partial sums) and one
// FFMA with an immediate -- no value table, no LDS.
template <bool FULL>
__device__ __forceinline__ void softmax_chunk_grid(const uint32_t (&r)[32], int nv, float x0, float y0, int brk4, float gw,
                                                   float c2, float& m, float& l, float& sx, float& sy) {
  float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float s = __uint_as_float(r[i]);
    cm[i & 3] = (FULL || i < nv) ? fmaxf(cm[i & 3], s) : cm[i & 3];
  }
  const float m_new = fmaxf(fmaxf(m, fmaxf(cm[0], cm[1])), fmaxf(cm[2], cm[3]));
  const float corr = ex2f((m - m_new) * c2);
  const float mb = m_new * c2;
  float g8[8], si[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = t * 4 + k;
      float e = ex2f(fmaf(__uint_as_float(r[i]), c2, -mb));
      if (!FULL) e = (i < nv) ? e : 0.f;
      acc += e;
      si[k] = fmaf(e, (float)i, si[k]);
    }
    g8[t] = acc;
  }
  float ls = 0.f, hs = 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    ls += g8[t];
    hs += (t >= brk4) ? g8[t] : 0.f;
  }
  const float sidx = (si[0] + si[1]) + (si[2] + si[3]);
  l = fmaf(l, corr, ls);
  sx = fmaf(sx, corr, fmaf(x0, ls, sidx) - gw * hs);
  sy = fmaf(sy, corr, fmaf(y0, ls, hs));
  m = m_new;
}
// Any grid width (small test grids): positions computed per element.
template <bool FULL>
__device__ __forceinline__ void softmax_chunk_grid_slow(const uint32_t (&r)[32], int nv, int col0, int gw, float c2, float& m,
                                                        float& l, float& sx, float& sy) {
  float cmax = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) cmax = (FULL || i < nv) ? fmaxf(cmax, __uint_as_float(r[i])) : cmax;
  const float m_new = fmaxf(m, cmax);
  const float corr = ex2f((m - m_new) * c2);
  const float mb = m_new * c2;
  float ls = 0.f, lx = 0.f, ly = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float e = ex2f(fmaf(__uint_as_float(r[i]), c2, -mb));
    if (!FULL) e = (i < nv) ? e : 0.f;
    const int c = col0 + i;
    ls += e;
    lx = fmaf(e, (float)(c % gw), lx);
    ly = fmaf(e, (float)(c / gw), ly);
  }
  l = fmaf(l, corr, ls);
  sx = fmaf(sx, corr, lx);
  sy = fmaf(sy, corr, ly);
  m = m_new;
}

// Emit 32 score columns of 32 rows (one warp) through the warp's smem stage and TMA bulk stores.
// WIDE = false: 2 KB stage, two [32 rows x 16 cols] boxes (64-byte swizzle); WIDE = true: 4 KB stage, one
// [32 rows x 32 cols] box (128-byte swizzle).  The swizzle is applied by hand so that the st.shared.v4 of the 32 rows
// are bank-conflict free.
template <bool WIDE>
__device__ __forceinline__ void emit_chunk_tma(const uint32_t (&r)[32], float scale, uint32_t stg_u32,
                                               const CUtensorMap* map, int col, int row0, int slab, int lane, int nk) {
  if constexpr (WIDE) {
    if (lane == 0) bulk_wait_read0();                 // the previous store has finished reading the stage
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 v;
      v.x = __uint_as_float(r[q * 4 + 0]) * scale;
      v.y = __uint_as_float(r[q * 4 + 1]) * scale;
      v.z = __uint_as_float(r[q * 4 + 2]) * scale;
      v.w = __uint_as_float(r[q * 4 + 3]) * scale;
      // CU_TENSOR_MAP_SWIZZLE_128B: 16-byte chunk index ^= address bits [7,9] = row & 7
      sts128(stg_u32 + lane * 128 + ((q ^ (lane & 7)) << 4), v);
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0 && col < nk) {                      // columns past nk inside the box are clipped by TMA
      tma_store_3d(map, stg_u32, col, row0, slab);
      bulk_commit();
    }
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 v;
        v.x = __uint_as_float(r[h * 16 + q * 4 + 0]) * scale;
        v.y = __uint_as_float(r[h * 16 + q * 4 + 1]) * scale;
        v.z = __uint_as_float(r[h * 16 + q * 4 + 2]) * scale;
        v.w = __uint_as_float(r[h * 16 + q * 4 + 3]) * scale;
        // CU_TENSOR_MAP_SWIZZLE_64B: 16-byte chunk index ^= address bits [7,8] = (row >> 1) & 3
        sts128(stg_u32 + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4), v);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0 && col + h * 16 < nk) {
        tma_store_3d(map, stg_u32, col + h * 16, row0, slab);
        bulk_commit();
      }
    }
  }
}

template <int NSMX, bool SK>
__global__ void __launch_bounds__((NSMX + 2) * 32, 1)
match_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                    const __grid_constant__ CUtensorMap map_s, KParams p) {
  constexpr int HALVES = NSMX / 8;                    // column halves of a key tile handled by different warps
  constexpr int STG_BYTES = R_BYTES / NSMX;           // emission stage per softmax warp
  constexpr int SMX_THREADS = NSMX * 32;
  constexpr int COLS_MINE = TN / HALVES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + OFF_BAR;
  auto q_full = [&](int g) { return bar0 + 8 * g; };
  const uint32_t q_empty = bar0 + 16;
  auto k_full = [&](int s) { return bar0 + 24 + 8 * s; };
  auto k_empty = [&](int s) { return bar0 + 24 + 8 * (STAGES + s); };
  auto s_full = [&](int g, int b) { return bar0 + 24 + 8 * (2 * STAGES + g * 2 + b); };
  auto s_empty = [&](int g, int b) { return bar0 + 24 + 8 * (2 * STAGES + 4 + g * 2 + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (p.nq + TM - 1) / TM;
  const int npair = (nqt + 1) / 2;
  const int nkt = (p.nk + TN - 1) / TN;
  const int n_items = p.nb * npair;
  const int nch = (p.terms == 3) ? NCHUNK : 2;        // single-pass mode streams the hi halves only
  // this CTA's span of (item, key tile) units, walked as visits (item, kt0, kt1)
  // SK: `u` runs over this CTA's span of units; else `u` is the item index, grid-strided
  const long long units = (long long)n_items * nkt;
  const int u_begin = SK ? (int)(units * blockIdx.x / gridDim.x) : (int)blockIdx.x;
  const int u_end = SK ? (int)(units * (blockIdx.x + 1) / gridDim.x) : n_items;
  auto next_visit = [&](int& u, int& item, int& kt0, int& kt1) {
    if constexpr (SK) {
      item = u / nkt;
      kt0 = u - item * nkt;
      kt1 = min(nkt, kt0 + (u_end - u));
      u += kt1 - kt0;
    } else {
      item = u;
      kt0 = 0;
      kt1 = nkt;
      u += (int)gridDim.x;
    }
  };

  if (threadIdx.x == 0) {
    mbar_init(q_full(0), 1);
    mbar_init(q_full(1), 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); }
    for (int g = 0; g < 2; ++g)
      for (int b = 0; b < 2; ++b) { mbar_init(s_full(g, b), 1); mbar_init(s_empty(g, b), 128 * HALVES); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NSMX + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + OFF_TMEM),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == NSMX) {
    // ===================== TMA producer =====================
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t kphase = 0;
      uint32_t it = 0;
      long long w_qe = 0, w_ke = 0;
      for (int u = u_begin; u < u_end; ++it) {
        int item, kt0, kt1;
        next_visit(u, item, kt0, kt1);
        const int prob = item / npair, qt0 = 2 * (item % npair);
        const int by = (prob + p.y_shift) % p.y_mod;
        w_qe += mbar_wait(q_empty, (it & 1) ^ 1);
        if (leader) {
          for (int g = 0; g < 2; ++g) {
            // a query tile past the end (odd tile count) is fully out of bounds: TMA fills it with zeros
            mbar_expect_tx(q_full(g), nch * CHUNK_BYTES);
            for (int c = 0; c < nch; ++c)
              tma_load_3d(sbase + OFF_Q + (g * NCHUNK + c) * CHUNK_BYTES, &map_x, q_full(g), c * CH_ELEMS,
                          (qt0 + g) * TM, prob);
          }
        }
        for (int kt = kt0; kt < kt1; ++kt) {
          for (int c = 0; c < nch; ++c) {
            w_ke += mbar_wait(k_empty(stage), kphase ^ 1);
            if (leader) {
              mbar_expect_tx(k_full(stage), CHUNK_BYTES);
              tma_load_3d(sbase + OFF_K + stage * CHUNK_BYTES, &map_y, k_full(stage), c * CH_ELEMS, kt * TN, by);
            }
            if (++stage == STAGES) { stage = 0; kphase ^= 1; }
          }
        }
      }
      if (p.prof && leader) { p.prof[blockIdx.x * 8 + 0] = w_qe; p.prof[blockIdx.x * 8 + 1] = w_ke; }
    }
    __syncwarp();
  } else if (warp == NSMX + 1) {
    // ===================== UMMA issuer =====================
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t kphase = 0;
      uint32_t it = 0, tile = 0;
      long long w_se = 0, w_kf = 0, w_qf = 0;
      const long long t_begin = clock64();
      const int n_tail = ((p.nk - (nkt - 1) * TN) + 15) & ~15;
      const uint32_t idesc_full = make_idesc(TN), idesc_tail = make_idesc(n_tail);
      // descriptor start-address units are 16 B: chunk = 1024 units, K16 step inside a swizzle row = 2 units
      const uint64_t qd = make_kmajor_sw128_desc(sbase + OFF_Q);
      // Notes from the role profile (tools/k1_roles.py): (1) tcgen05.mma issue back-pressures, the issuer is busy for
      // the whole UMMA time; (2) hoisting the mbarrier waits of the next chunk in front of the last UMMAs of the
      // current one made the kernel SLOWER (r1n: 97 -> 112 us): a try_wait on a barrier that is not complete yet
      // suspends the warp with a coarse wake-up and the tensor pipe drains meanwhile -- waits stay between chunks;
      // (3) the UMMAs of the two query tiles are interleaved one by one so that consecutive instructions
      // accumulate into different TMEM tiles.
      for (int u = u_begin; u < u_end; ++it) {
        int item, kt0, kt1;
        next_visit(u, item, kt0, kt1);
        for (int kt = kt0; kt < kt1; ++kt, ++tile) {
          const int buf = tile & 1;
          const uint32_t use = tile >> 1;
          const uint32_t idesc = (kt == nkt - 1) ? idesc_tail : idesc_full;
          for (int c = 0; c < nch; ++c) {
            w_kf += mbar_wait(k_full(stage), kphase);
            if (c == 0) {
              // the accumulators of this key step must have been drained by the softmax warps (two steps ago)
              w_se += mbar_wait(s_empty(0, buf), (use & 1) ^ 1);
              w_se += mbar_wait(s_empty(1, buf), (use & 1) ^ 1);
              if (kt == kt0) {
                w_qf += mbar_wait(q_full(0), it & 1);
                w_qf += mbar_wait(q_full(1), it & 1);
              }
            }
            tc_fence_after();
            if (leader) {
              const uint64_t kd = make_kmajor_sw128_desc(sbase + OFF_K + stage * CHUNK_BYTES);
              const uint32_t d0 = tmem_base + (uint32_t)(buf * TN), d1 = d0 + 2 * TN;
              // chunk 0/1 = Y.hi halves: pair with X.hi and X.lo ; chunk 2/3 = Y.lo halves: pair with X.hi
              const uint64_t a_hi0 = qd + (uint64_t)((c & 1) * (CHUNK_BYTES >> 4));
              const uint64_t a_hi1 = a_hi0 + (uint64_t)(NCHUNK * (CHUNK_BYTES >> 4));
              const uint32_t acc0 = c ? 1u : 0u;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(d0, a_hi0 + 2 * k, kd + 2 * k, idesc, (k ? 1u : acc0));
                umma_bf16(d1, a_hi1 + 2 * k, kd + 2 * k, idesc, (k ? 1u : acc0));
              }
              if (c < 2 && nch == NCHUNK) {
                const uint64_t a_lo0 = a_hi0 + (uint64_t)(2 * (CHUNK_BYTES >> 4)), a_lo1 = a_hi1 + (uint64_t)(2 * (CHUNK_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(d0, a_lo0 + 2 * k, kd + 2 * k, idesc, 1u);
                  umma_bf16(d1, a_lo1 + 2 * k, kd + 2 * k, idesc, 1u);
                }
              }
              umma_commit(k_empty(stage));       // frees the smem stage when these MMAs retire
              if (c == nch - 1) {                // both S tiles complete -> softmax groups
                umma_commit(s_full(0, buf));
                umma_commit(s_full(1, buf));
              }
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; kphase ^= 1; }
          }
        }
        if (leader) umma_commit(q_empty);        // query tiles no longer read -> producer may overwrite
        __syncwarp();
      }
      if (p.prof && leader) {
        p.prof[blockIdx.x * 8 + 2] = w_se; p.prof[blockIdx.x * 8 + 3] = w_kf; p.prof[blockIdx.x * 8 + 4] = w_qf;
        p.prof[blockIdx.x * 8 + 5] = clock64() - t_begin;
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax warps =====================
    const int g = warp / (4 * HALVES);                   // query tile of the pair
    const int half = (warp >> 2) & (HALVES - 1);         // column half of every key tile (HALVES == 2)
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
    const int r_in_tile = quarter * 32 + lane;
    const int st = threadIdx.x;                          // index within the softmax warps
    const bool grid_mode = p.grid_w > 0;
    const bool grid_fast = p.grid_w >= 32 && (p.grid_w & 3) == 0;
    const float gwf = (float)p.grid_w;
    float* vtab = reinterpret_cast<float*>(smem + OFF_V);
    const uint32_t stg = sbase + OFF_STG + warp * STG_BYTES;
    const uint32_t vtab_u32 = sbase + OFF_V;
    // end-of-item merge slot (16 B per row): the warp's own emission stage in matching mode, a dedicated slot otherwise
    const uint32_t my_slot = grid_mode ? stg : sbase + OFF_MERGE + warp * 512;
    const uint32_t partner_slot = grid_mode ? stg + 4 * STG_BYTES : sbase + OFF_MERGE + (warp + 4) * 512;
    const float c2 = 1.4426950408889634f * p.inv_sqrt_c;   // log2(e)/sqrt(C)
    const int nk_pad = (p.nk + 31) & ~31;
    const int cbeg = half * COLS_MINE;
    uint32_t tile = 0;
    int cur_v = -1;
    long long w_sf = 0;
    const long long t_begin = clock64();
    for (int u = u_begin; u < u_end;) {
      int item, kt0, kt1;
      next_visit(u, item, kt0, kt1);
      const int prob = item / npair, qt = 2 * (item % npair) + g;
      const int row = qt * TM + r_in_tile;
      if (!grid_mode) {
        const int v_id = (p.v_stride_b == 0) ? 0 : prob;
        if (v_id != cur_v) {
          asm volatile("bar.sync 1, %0;" ::"n"(SMX_THREADS) : "memory");   // everyone finished with the old table
          const float* vg = p.v + (size_t)prob * p.v_stride_b;
          // masked tail columns are multiplied by p = 0: keep them finite
          for (int c = st; c < nk_pad; c += SMX_THREADS) {
            vtab[c] = (c < p.nk) ? __ldg(vg + c) : 0.f;
            vtab[MAXK + c] = (c < p.nk) ? __ldg(vg + p.nk + c) : 0.f;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(SMX_THREADS) : "memory");
          cur_v = v_id;
        }
      }
      const int row0w = qt * TM + quarter * 32;          // first row of this warp's 32-row slab
      const bool emit = p.s_mode != 0 && prob >= p.s_first && prob < p.s_first + p.s_count && row0w < p.nq;
      const int slab = prob - p.s_first;
      // grid position of this warp's first column of the current key tile (matching mode)
      int gx = grid_mode ? (kt0 * TN + cbeg) % p.grid_w : 0, gy = grid_mode ? (kt0 * TN + cbeg) / p.grid_w : 0;
      float m = -INFINITY, l = 0.f, sx = 0.f, sy = 0.f;
      for (int kt = kt0; kt < kt1; ++kt, ++tile) {
        const int buf = tile & 1;
        const uint32_t use = tile >> 1;
        w_sf += mbar_wait(s_full(g, buf), use & 1);
        tc_fence_after();
        const int col_base = kt * TN;
        const int n_valid = min(TN, p.nk - col_base);
        const int n_mine = max(0, min(COLS_MINE, n_valid - cbeg));
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((g * 2 + buf) * TN + cbeg);
        auto work = [&](const uint32_t (&r)[32], int c0, int nv, bool full) {      // c0 = column offset within my part
          const int col0 = col_base + cbeg + c0;
          if (emit) {
            if (p.s_mode == 1) {
              emit_chunk_tma<HALVES == 1>(r, p.inv_sqrt_c, stg, &map_s, col0, row0w, slab, lane, p.nk);
            } else if (row < p.nq) {
              float* sg = p.s_out + ((size_t)slab * p.nq + row) * p.nk + col0;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < nv) sg[i] = __uint_as_float(r[i]) * p.inv_sqrt_c;
            }
          }
          if (grid_fast) {
            int x0 = gx + c0, y0 = gy;
            while (x0 >= p.grid_w) { x0 -= p.grid_w; ++y0; }
            const int rem = p.grid_w - x0;                 // columns left in this grid row
            const int brk4 = rem < 32 ? (rem >> 2) : 8;
            if (full) softmax_chunk_grid<true>(r, 32, (float)x0, (float)y0, brk4, gwf, c2, m, l, sx, sy);
            else softmax_chunk_grid<false>(r, nv, (float)x0, (float)y0, brk4, gwf, c2, m, l, sx, sy);
          } else if (grid_mode) {
            if (full) softmax_chunk_grid_slow<true>(r, 32, col0, p.grid_w, c2, m, l, sx, sy);
            else softmax_chunk_grid_slow<false>(r, nv, col0, p.grid_w, c2, m, l, sx, sy);
          } else {
            const uint32_t vx = vtab_u32 + 4 * col0;
            if (full) softmax_chunk<true>(r, 32, vx, vx + 4 * MAXK, c2, m, l, sx, sy);
            else softmax_chunk<false>(r, nv, vx, vx + 4 * MAXK, c2, m, l, sx, sy);
          }
        };
        if (n_mine == COLS_MINE) {
          if constexpr (HALVES == 1) {
            // 8 softmax warps: software pipeline, the TMEM load of chunk c+1 is in flight while chunk c is processed
            uint32_t ra[32], rb[32];
            tmem_ld32_async(taddr, ra);
            tmem_wait(ra);
            tmem_ld32_async(taddr + 32, rb);
            work(ra, 0, 32, true);
            tmem_wait(rb);
            tmem_ld32_async(taddr + 64, ra);
            work(rb, 32, 32, true);
            tmem_wait(ra);
            tmem_ld32_async(taddr + 96, rb);
            work(ra, 64, 32, true);
            tmem_wait(rb);
            tc_fence_before();
            mbar_arrive(s_empty(g, buf));                // accumulator drained: the next MMAs may overwrite it
            work(rb, 96, 32, true);
          } else {
            // 16 softmax warps leave 96 registers per thread: one buffer, the other warps hide the TMEM latency
            uint32_t ra[32];
            tmem_ld32_async(taddr, ra);
            tmem_wait(ra);
            work(ra, 0, 32, true);
            tmem_ld32_async(taddr + 32, ra);
            tmem_wait(ra);
            tc_fence_before();
            mbar_arrive(s_empty(g, buf));
            work(ra, 32, 32, true);
          }
        } else {
          for (int c0 = 0; c0 < n_mine; c0 += 32) {
            uint32_t ra[32];
            tmem_ld32_async(taddr + c0, ra);
            tmem_wait(ra);
            const int nv = min(32, n_mine - c0);
            work(ra, c0, nv, nv == 32);
          }
          tc_fence_before();
          mbar_arrive(s_empty(g, buf));
        }
        if (grid_fast) {                                 // next key tile: + TN columns
          gx += TN;
          while (gx >= p.grid_w) { gx -= p.grid_w; ++gy; }
        }
      }
      if constexpr (HALVES == 2) {
        // merge the two column halves of every row: half 1 hands (m, l, sx, sy) to half 0 through shared memory
        if (half == 1) {
          if (p.s_mode == 1) {
            if (lane == 0) bulk_wait_read0();            // my emission stage is free again
            __syncwarp();
          }
          sts128(my_slot + lane * 16, make_float4(m, l, sx, sy));
        }
        asm volatile("bar.sync %0, 256;" ::"r"(2 + g) : "memory");
        if (half == 0) {
          const float4 o = lds128(partner_slot + lane * 16);
          const float mm = fmaxf(m, o.x);
          const float a = ex2f((m - mm) * c2), b = ex2f((o.x - mm) * c2);   // an empty half has m = -inf, l = 0
          l = l * a + o.y * b;
          sx = sx * a + o.z * b;
          sy = sy * a + o.w * b;
          m = mm;
        }
      }
      if (SK && (kt0 != 0 || kt1 != nkt)) {
        // This CTA saw only part of the item's keys.  The CTAs covering the item are consecutive: first .. last.
        const long long ub = (long long)item * nkt;
        const int first = (int)(((ub + 1) * gridDim.x + units - 1) / units) - 1;
        const int last = (int)(((ub + nkt) * gridDim.x + units - 1) / units) - 1;
        const int parts = last - first + 1;
        float4* slab = p.part + (size_t)item * p.pmax * (2 * TM);
        if (half == 0) slab[((int)blockIdx.x - first) * (2 * TM) + g * TM + r_in_tile] = make_float4(m, l, sx, sy);
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(SMX_THREADS) : "memory");
        volatile uint32_t* flag = tmem_slot + 2;
        if (st == 0) {
          const bool is_last = atomicAdd(p.cnt + item, 1u) == (unsigned)(parts - 1);
          if (is_last) p.cnt[item] = 0;                  // left at zero for the next launch
          *flag = is_last ? 1u : 0u;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(SMX_THREADS) : "memory");
        if (*flag == 0u) {
          l = 0.f;                                       // somebody else finishes these rows
        } else {
          __threadfence();
          // merge in part order, whoever arrives last: the result does not depend on the schedule
          m = -INFINITY; l = 0.f; sx = 0.f; sy = 0.f;
          if (half == 0)
            for (int q = 0; q < parts; ++q) {
              const float4 o = __ldcg(slab + q * (2 * TM) + g * TM + r_in_tile);
              const float mm = fmaxf(m, o.x);
              const float a = ex2f((m - mm) * c2), b = ex2f((o.x - mm) * c2);
              l = l * a + o.y * b;
              sx = sx * a + o.z * b;
              sy = sy * a + o.w * b;
              m = mm;
            }
        }
      }
      if (half == 0 && row < p.nq && (!SK || l > 0.f)) {
        float ex = sx / l, ey = sy / l;
        if (p.sub_grid) { ex -= (float)(row % p.grid_w); ey -= (float)(row / p.grid_w); }
        p.out[((size_t)prob * 2 + 0) * p.nq + row] = ex;
        p.out[((size_t)prob * 2 + 1) * p.nq + row] = ey;
        if (p.lse != nullptr) p.lse[(size_t)prob * p.nq + row] = m * p.inv_sqrt_c + logf(l);
      }
      if constexpr (HALVES == 2) asm volatile("bar.sync %0, 256;" ::"r"(2 + g) : "memory");   // merge slot reusable
    }
    if (p.prof && threadIdx.x == 0) { p.prof[blockIdx.x * 8 + 6] = w_sf; p.prof[blockIdx.x * 8 + 7] = clock64() - t_begin; }
    if (lane == 0) bulk_wait0();                         // all score stores of this warp have landed
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NSMX + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---- operand split: fp32 -> bf16 hi | bf16 lo, token-major -------------------
__device__ __forceinline__ void split2(float a, float b, __nv_bfloat162& hi, __nv_bfloat162& lo) {
  __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  hi = __halves2bfloat162(ah, bh);
  lo = __halves2bfloat162(__float2bfloat16_rn(a - __bfloat162float(ah)), __float2bfloat16_rn(b - __bfloat162float(bh)));
}

// src [nb][128][n] (channel-major; batches >= nb come from src2) -> dst [..][n][256]
// One thread = one token x 32 channels: 32 coalesced 4-byte loads (lanes = consecutive tokens), then the token's
// 64-byte hi and lo segments as 4 + 4 128-bit stores (full 32-byte sectors).  No shared memory, no barrier:
// 62 MB of traffic at c2 in ~10 us instead of 22 us for the smem-transpose version.
__global__ void __launch_bounds__(128)
split_cn_kernel(const float* __restrict__ src, const float* __restrict__ src2, __nv_bfloat16* __restrict__ dst, int nb,
                int n) {
  const int b = blockIdx.z, chunk = blockIdx.y;
  const int tok = blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= n) return;
  const float* s = (b < nb ? src + (size_t)b * 128 * n : src2 + (size_t)(b - nb) * 128 * n) + (size_t)chunk * 32 * n + tok;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __ldg(s + (size_t)i * n);
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    __nv_bfloat162 h, l;
    split2(v[2 * i], v[2 * i + 1], h, l);
    hi[i] = *reinterpret_cast<uint32_t*>(&h);
    lo[i] = *reinterpret_cast<uint32_t*>(&l);
  }
  uint4* d = reinterpret_cast<uint4*>(dst + ((size_t)b * n + tok) * 256 + chunk * 32);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    d[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
    d[16 + i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);     // + 128 bf16 = 256 bytes
  }
}

// src [rows][128] (token-major) -> dst [rows][256]
__global__ void __launch_bounds__(256)
split_nc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one float2 per thread
  if (i >= rows * 64) return;
  long long r = i >> 6;
  int cp = (int)(i & 63);
  float2 v = __ldg(reinterpret_cast<const float2*>(src + r * 128) + cp);
  __nv_bfloat162 hi, lo;
  split2(v.x, v.y, hi, lo);
  *reinterpret_cast<__nv_bfloat162*>(dst + r * 256 + 2 * cp) = hi;
  *reinterpret_cast<__nv_bfloat162*>(dst + r * 256 + 128 + 2 * cp) = lo;
}

// ---- host side -----------------------------------------------------------------
// [nb][n][256] bf16, box = 64 elements x 128 rows x 1 batch, 128-byte swizzle, OOB rows read as zero
int make_map(CUtensorMap* m, const void* base, int nb, int n) {
  EncodeTiledFn enc = get_encoder();
  if (enc == nullptr) {
    emip_set_error("cuTensorMapEncodeTiled not available from the driver");
    return EMIP_ENOSYS;
  }
  cuuint64_t dims[3] = {256, (cuuint64_t)n, (cuuint64_t)nb};
  cuuint64_t strides[2] = {512, (cuuint64_t)n * 512};
  cuuint32_t box[3] = {CH_ELEMS, TM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    emip_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return EMIP_EINVAL;
  }
  return EMIP_OK;
}

// scores [slabs][nq][nk] fp32, box = 16 (64-byte swizzle) or 32 (128-byte swizzle) columns x 32 rows x 1 slab
// (store side: OOB is clipped)
int make_store_map(CUtensorMap* m, float* base, int slabs, int nq, int nk, bool wide) {
  EncodeTiledFn enc = get_encoder();
  if (enc == nullptr) return EMIP_ENOSYS;
  cuuint64_t dims[3] = {(cuuint64_t)nk, (cuuint64_t)nq, (cuuint64_t)slabs};
  cuuint64_t strides[2] = {(cuuint64_t)nk * 4, (cuuint64_t)nq * nk * 4};
  cuuint32_t box[3] = {wide ? 32u : 16u, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EMIP_OK : EMIP_EINVAL;
}

}  // namespace

static unsigned long long* g_prof = nullptr;
static int g_softmax_warps = 0;
// Diagnostics: force 8 or 16 softmax warps per CTA (tools/k1_roles.py compares them); 0 = choose by mode:
// 8 for the 3-term split (UMMA-bound: fewer warps = no spills, less smem/L1 traffic next to the operand reads),
// 16 for single-pass bf16 (softmax-bound: r1p 70 vs 76 us).
extern "C" void emip_match_tc_set_variant(int softmax_warps) {
  g_softmax_warps = (softmax_warps == 16 || softmax_warps == 8) ? softmax_warps : 0;
}

// Diagnostics: device buffer of gridDim.x * 8 counters filled by the next launches (NULL switches it off).
extern "C" void emip_match_tc_set_profile_buffer(unsigned long long* dev_buf) { g_prof = dev_buf; }

bool match_tc_supported(int nq, int nk, int c) { return c == 128 && nk <= MAXK && nk >= 16 && nq >= 1; }

size_t match_tc_split_bytes(int nb, int n, int c) { return emip_align_up((size_t)nb * n * 2 * c * 2, 1024); }

// Upper bound of the stream-K workspace of match_tc_fwd for nb problems: with G CTAs and T = items * key tiles units,
// a span holds >= floor(T / G) units, so an item (nkt units) meets at most nkt / floor(T / G) + 2 spans, and never
// more than nkt.
size_t match_tc_streamk_bytes(int nb, int nq, int nk) {
  const int nqt = (nq + TM - 1) / TM, nkt = (nk + TN - 1) / TN;
  const long long n_items = (long long)nb * ((nqt + 1) / 2), units = n_items * nkt;
  if (units == 0) return 0;
  const long long g = units < emip_num_sms() ? units : emip_num_sms();
  long long pmax = nkt / (units / g) + 2;
  if (pmax > nkt) pmax = nkt;
  return emip_align_up(emip_align_up((size_t)n_items * sizeof(unsigned), 256) + (size_t)n_items * pmax * 2 * TM * sizeof(float4), 1024);
}

int match_tc_split(const float* src, const float* src2, void* dst, int nb, int n, int c, int layout, int dst_batch0,
                   cudaStream_t st) {
  if (c != 128) { emip_set_error("match_tc_split: C=%d unsupported", c); return EMIP_ENOSYS; }
  if (nb == 0 || n == 0) return EMIP_OK;
  __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dst) + (size_t)dst_batch0 * n * 256;
  if (layout == EMIP_LAYOUT_CN) {
    dim3 grid((n + 127) / 128, 4, src2 ? 2 * nb : nb);
    split_cn_kernel<<<grid, 128, 0, st>>>(src, src2, d, nb, n);
  } else {
    long long rows = (long long)nb * n;
    split_nc_kernel<<<(unsigned)((rows * 64 + 255) / 256), 256, 0, st>>>(src, d, rows);
    if (src2) split_nc_kernel<<<(unsigned)((rows * 64 + 255) / 256), 256, 0, st>>>(src2, d + rows * 256, rows);
  }
  EMIP_CHECK_LAUNCH("match_tc_split");
  return EMIP_OK;
}

int match_tc_fwd(const MatchTcArgs& a, cudaStream_t st) {
  if (a.nb == 0) return EMIP_OK;
  // the value table (attention mode) holds MAXK columns; the analytic grid (matching mode) has no such limit
  if (a.nq < 1 || a.nk < 16 || (a.grid_w <= 0 && a.nk > MAXK)) { emip_set_error("match_tc_fwd: unsupported shape"); return EMIP_ENOSYS; }
  if (a.terms != 1 && a.terms != 3) { emip_set_error("match_tc_fwd: terms must be 1 or 3"); return EMIP_EINVAL; }
  if (a.sub_grid && a.grid_w <= 0) { emip_set_error("match_tc_fwd: sub_grid needs grid_w"); return EMIP_EINVAL; }
  CUtensorMap mx, my, ms;
  int rc;
  const int nsmx = g_softmax_warps ? g_softmax_warps : (a.terms == 1 ? 16 : 8);
  if ((rc = make_map(&mx, a.x_split, a.nbx, a.nq))) return rc;
  if ((rc = make_map(&my, a.y_split, a.nby, a.nk))) return rc;
  ms = mx;
  int s_mode = 0;
  if (a.s_out != nullptr && a.s_count > 0) {
    // TMA needs 16-byte aligned global rows; odd token counts fall back to plain stores
    const bool tma_ok = (a.nk % 4 == 0) && (reinterpret_cast<uintptr_t>(a.s_out) % 16 == 0) &&
                        make_store_map(&ms, a.s_out, a.s_count, a.nq, a.nk, nsmx == 8) == EMIP_OK;
    s_mode = tma_ok ? 1 : 2;
    if (!tma_ok) ms = mx;
  }
  if (int rc__ = emip_func_max_smem((const void*)(match_tc_fwd_kernel<8, false>), SMEM_BYTES)) return rc__;
  if (int rc__ = emip_func_max_smem((const void*)(match_tc_fwd_kernel<16, false>), SMEM_BYTES)) return rc__;
  if (int rc__ = emip_func_max_smem((const void*)(match_tc_fwd_kernel<8, true>), SMEM_BYTES)) return rc__;
  if (int rc__ = emip_func_max_smem((const void*)(match_tc_fwd_kernel<16, true>), SMEM_BYTES)) return rc__;
  KParams p;
  p.v = a.v; p.v_stride_b = a.v_stride_b; p.out = a.out; p.lse = a.lse; p.s_out = a.s_out;
  p.nb = a.nb; p.nq = a.nq; p.nk = a.nk; p.y_shift = a.y_shift; p.y_mod = a.y_mod;
  p.s_first = a.s_first; p.s_count = a.s_count;
  p.grid_w = a.grid_w; p.sub_grid = a.sub_grid; p.terms = a.terms; p.s_mode = s_mode;
  p.inv_sqrt_c = 1.0f / a.sqrt_c;
  p.prof = g_prof;
  const int nqt = (a.nq + TM - 1) / TM, nkt = (a.nk + TN - 1) / TN;
  const int n_items = a.nb * ((nqt + 1) / 2);
  const long long units = (long long)n_items * nkt;
  // Whole items grid-strided keep every SM busy when the batch is large; when the last wave would leave more than a
  // fifth of the SM-time idle (small batches: B = 1 has 16 items for 148 SMs) the units are dealt in equal spans.
  // (At c2, 256 items, both schedules take the same time: the kernel runs power-limited at ~1.5 GHz when all SMs work,
  // and the score-emitting items -- the second half of the list -- unbalance contiguous spans.)
  const int sms = emip_num_sms();
  const int waves = (n_items + sms - 1) / sms;
  const int stream_k = a.schedule == 1 ? 1 : a.schedule == 2 ? 0 : ((double)n_items < 0.8 * waves * sms ? 1 : 0);
  const int grid = stream_k ? (int)(units < sms ? units : sms) : (n_items < sms ? n_items : sms);
  p.cnt = nullptr; p.part = nullptr; p.pmax = 1;
  if (stream_k) {
    // bookkeeping (arrival counters + partial rows) carved from the caller's workspace
    const size_t cnt_bytes = emip_align_up((size_t)n_items * sizeof(unsigned), 256);
    int pmax = 1;
    for (int i = 0; i < n_items; ++i) {
      const long long ub = (long long)i * nkt;
      const int first = (int)(((ub + 1) * grid + units - 1) / units) - 1, last = (int)(((ub + nkt) * grid + units - 1) / units) - 1;
      if (last - first + 1 > pmax) pmax = last - first + 1;
    }
    const size_t need = cnt_bytes + (size_t)n_items * pmax * 2 * TM * sizeof(float4);
    if (a.sk_ws == nullptr || a.sk_bytes < need || reinterpret_cast<uintptr_t>(a.sk_ws) % 16 != 0) {
      emip_set_error("match_tc_fwd: stream-K workspace too small or misaligned (%zu < %zu bytes)", a.sk_bytes, need);
      return EMIP_ENOMEM;
    }
    p.cnt = static_cast<unsigned*>(a.sk_ws);
    p.part = reinterpret_cast<float4*>(static_cast<char*>(a.sk_ws) + cnt_bytes);
    p.pmax = pmax;
  }
  if (stream_k) EMIP_CUDA(cudaMemsetAsync(p.cnt, 0, (size_t)n_items * sizeof(unsigned), st));
  if (nsmx == 8) {
    if (stream_k) match_tc_fwd_kernel<8, true><<<grid, 10 * 32, SMEM_BYTES, st>>>(mx, my, ms, p);
    else match_tc_fwd_kernel<8, false><<<grid, 10 * 32, SMEM_BYTES, st>>>(mx, my, ms, p);
  } else {
    if (stream_k) match_tc_fwd_kernel<16, true><<<grid, 18 * 32, SMEM_BYTES, st>>>(mx, my, ms, p);
    else match_tc_fwd_kernel<16, false><<<grid, 18 * 32, SMEM_BYTES, st>>>(mx, my, ms, p);
  }
  EMIP_CHECK_LAUNCH("match_tc_fwd");
  return EMIP_OK;
}
