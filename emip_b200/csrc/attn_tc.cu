// Fused single-pass attention forward on the 5th-gen tensor cores:
//   out = softmax(Q K^T / sqrt(C)) V,   C = d_v = 128, fp32 in / fp32 out
// for a5 (EMIP_long memory read, model/EMIP_long/LTM.py:49-68: queries = the current frame's key map, keys / values =
// the space-time memory) and f2 (the attention core of the GMFlow FeatureTransformer,
// model/EMIP_short/motion/gmflow/transformer.py:8-16 and the per-window attention of :46-105).
//
// One kernel, flash style: the score tile S lives in TMEM, the probabilities go back to TMEM as the A operand of the
// second UMMA, the N x M score matrix never exists.  It replaces the two-pass scheme of r1 (row log-sum-exp by the
// matching kernel, then e^{S-L} V by the gradient kernel: three S-sized MMA passes and ~8 launches).
//
// CTA = one 128-row query tile, persistent over (problem, row tile, key split) items, 576 threads:
//   warps 0..15 softmax + epilogue (thread <-> TMEM lane <-> query row; warp w: lane quarter w%4, 32-column part w/4 of
//               every key tile; the four threads of a row exchange their tile maxima through shared memory)
//   warp 16     TMA producer: Q tile (token-major hi|lo) once per item, then per 128-key tile four token-major chunks
//               of K (for S) and four token-major chunks of V (for P V) through one 8-stage ring
//   warp 17     UMMA issuer
// TMEM (512 columns): S double-buffered [0,256) | P as bf16 hi [256,320) + lo [320,384) | O accumulator [384,512).
//   UMMA-1 (SS): S  = Q.hi K.hi^T + Q.lo K.hi^T + Q.hi K.lo^T        (3-term bf16 split: fp32-grade scores)
//   UMMA-2 (TS): O += P.hi V.hi + P.lo V.hi + P.hi V.lo              (A = P read from TMEM, B = V tiles as they lie in
//                memory: token-major V through an MN-major shared-memory descriptor, channel-major V (the memory read)
//                through the K-major one -- either way the operand-split pass of V is elementwise, no transposed copy)
// Online softmax with a lazy reference: P = 2^{(S - ref)} with ref = the running row maximum as of the last time it
// grew by more than 2^TAU; only then is the O accumulator rescaled (tcgen05.ld / multiply / tcgen05.st between two
// UMMA-2 groups).  Any ref gives the same out = O / l in exact arithmetic; bf16 hi|lo and the fp32 accumulators have the
// fp32 exponent range, so P <= 2^TAU costs no precision.  The issuer runs UMMA-1 of tile t+1 ahead of UMMA-2 of tile t.
#include "common.cuh"
#include "pair_common.cuh"
#include "tc_common.cuh"
#include "attn_tc.cuh"

namespace {
using namespace tc;

constexpr int TN = 128;                 // key columns per tile
constexpr int CH_ELEMS = 64;
constexpr int CHUNK_BYTES = TM * 128;   // 16 KB
constexpr int STAGES = 8;
static_assert(STAGES % 4 == 0, "the MN-major V operand needs its two 64-channel chunks in adjacent ring stages");
constexpr int NMATH = 16;               // softmax warps
constexpr int NTHREADS = (NMATH + 2) * 32;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_P = 256, COL_O = 384;
constexpr float TAU = 16.0f;            // log2 units: rescale only when the row maximum grew by more than 2^16

constexpr int OFF_Q = 0;                                  // 4 chunks: hi[0:64] hi[64:128] lo[0:64] lo[64:128]
constexpr int OFF_RING = OFF_Q + 4 * CHUNK_BYTES;
constexpr int OFF_XM = OFF_RING + STAGES * CHUNK_BYTES;   // [2 parities][16 warps][32 lanes] tile maxima
constexpr int OFF_XL = OFF_XM + 2 * NMATH * 32 * 4;       // [16 warps][32 lanes] partial row sums (epilogue)
constexpr int OFF_BAR = OFF_XL + NMATH * 32 * 4;
constexpr int NBAR = 2 + 2 * STAGES + 4 + 4;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

constexpr int MAXSEG = 3;
struct AParams {
  AttnTcArgs a;
  float inv_sqrt_c;
  unsigned long long* prof;   // optional [gridDim.x][16] cycle counters (emip_attn_tc_set_profile_buffer)
  // problem sets of the launch (set 0 = the main arguments): items of set s are [sum of seg_items[< s], + seg_items[s])
  int nseg, seg_items[MAXSEG], seg_nq[MAXSEG], seg_nk[MAXSEG], seg_prob0[MAXSEG];      // seg_prob0 = blk0 * win.B
};
struct AMaps { CUtensorMap q[MAXSEG], k[MAXSEG], v[MAXSEG]; };

// item -> (set, problem, row tile, key split) and the set's sizes
struct AItem {
  int seg, prob, rt, ks, nq, nk, nkt, kb, ke;
};
__device__ __forceinline__ AItem locate_item(const AParams& ap, int item, int ns) {
  AItem r;
  r.seg = 0;
  while (r.seg + 1 < ap.nseg && item >= ap.seg_items[r.seg]) { item -= ap.seg_items[r.seg]; ++r.seg; }
  r.nq = ap.seg_nq[r.seg];
  r.nk = ap.seg_nk[r.seg];
  const int nrt = (r.nq + TM - 1) / TM;
  r.nkt = (r.nk + TN - 1) / TN;
  r.ks = item % ns;
  r.rt = (item / ns) % nrt;
  r.prob = item / (ns * nrt);
  r.kb = r.ks * r.nkt / ns;
  r.ke = (r.ks + 1) * r.nkt / ns;
  return r;
}

__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

template <bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ AMaps maps, const __grid_constant__ AParams ap) {
  const AttnTcArgs& p = ap.a;
  pdl_trigger();                                          // common.cuh: the next kernel's prologue may overlap this kernel's tail
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + OFF_BAR;
  const uint32_t q_full = bar0, q_empty = bar0 + 8;
  auto r_full = [&](int s) { return bar0 + 16 + 8 * s; };
  auto r_empty = [&](int s) { return bar0 + 16 + 8 * (STAGES + s); };
  auto s_full = [&](int b) { return bar0 + 16 + 8 * (2 * STAGES + b); };
  auto s_empty = [&](int b) { return bar0 + 16 + 8 * (2 * STAGES + 2 + b); };
  const uint32_t p_full = bar0 + 16 + 8 * (2 * STAGES + 4), p_empty = p_full + 8, o_full = p_full + 16, o_empty = p_full + 24;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // an item is (problem set, problem, row tile, key split): split ks covers key tiles [ks*nkt/ns, (ks+1)*nkt/ns)
  const int ns = p.ksplit > 1 ? p.ksplit : 1;
  int n_items = 0;
  for (int s = 0; s < ap.nseg; ++s) n_items += ap.seg_items[s];

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(r_full(s), 1); mbar_init(r_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(s_empty(b), NMATH * 32); }
    mbar_init(p_full, NMATH * 32);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, NMATH * 32);
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
  pdl_wait();                                             // everything below reads what the kernel in front of this one wrote

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
      const AItem w = locate_item(ap, item, ns);
      const int rt = w.rt, prob = w.prob, kb = w.kb, ke = w.ke;
      const CUtensorMap* map_q = &maps.q[w.seg];
      const CUtensorMap* map_k = &maps.k[w.seg];
      const CUtensorMap* map_v = &maps.v[w.seg];
      mbar_wait(q_empty, (it & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(q_full, 4 * CHUNK_BYTES);
        for (int c = 0; c < 4; ++c) tma_load_3d(sbase + OFF_Q + c * CHUNK_BYTES, map_q, q_full, c * CH_ELEMS, rt * TM, prob);
      }
      // ring order = consumption order of the issuer: K(0), then per tile { K(t+1), V(t) }
      for (int c = 0; c < 4; ++c) push(map_k, c * CH_ELEMS, kb * TN, prob);
      for (int t = kb; t < ke; ++t) {
        if (t + 1 < ke)
          for (int c = 0; c < 4; ++c) push(map_k, c * CH_ELEMS, (t + 1) * TN, prob);
        if (p.v_chn)
          for (int c = 0; c < 4; ++c)     // channel-major V: hi keys[0:64], hi keys[64:128], lo keys[0:64], lo keys[64:128]
            push(map_v, t * TN + (c & 1) * CH_ELEMS, (c >> 1) * 128, prob);
        else
          for (int c = 0; c < 4; ++c)     // token-major V like K: hi ch[0:64], hi ch[64:128], lo ch[0:64], lo ch[64:128]
            push(map_v, c * CH_ELEMS, t * TN, prob);
      }
    }
    __syncwarp();
  } else if (warp == NMATH + 1) {
    // ===================== UMMA issuer =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0, it = 0, tile = 0;
    const uint32_t idesc_full = make_idesc(TN);
    uint32_t idesc_tail = idesc_full;                      // per item: the set's last key tile, rounded up to 16 columns
    const uint64_t qd = make_kmajor_sw128_desc(sbase + OFF_Q);
    const uint32_t t_p = tmem_base + COL_P, t_o = tmem_base + COL_O;
    long long w_se = 0, w_rf = 0, w_pf = 0, w_oe = 0, w_qf = 0;
    const long long t_begin = PROF ? clock64() : 0;
    auto mma1 = [&](uint32_t T, bool tail) {              // S tile T -> TMEM buffer T & 1
      const int buf = T & 1;
      prof_add<PROF>(w_se, mbar_wait(s_empty(buf), ((T >> 1) & 1) ^ 1));
      const uint32_t idesc = tail ? idesc_tail : idesc_full;
      const uint32_t d = tmem_base + (uint32_t)(buf * TN);
      for (int c = 0; c < 4; ++c) {
        prof_add<PROF>(w_rf, mbar_wait(r_full(stage), phase));
        tc_fence_after();
        if (leader) {
          const uint64_t kd = make_kmajor_sw128_desc(sbase + OFF_RING + stage * CHUNK_BYTES);
          const uint64_t a_hi = qd + (uint64_t)((c & 1) * (CHUNK_BYTES >> 4));
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
    // UMMA-2 reads V as an MN-major B operand straight from the token-major tile (keys = rows of 128 B, 64 channels each):
    // the two 64-channel chunks of a half are adjacent ring stages (the ring advances in groups of four chunks and
    // STAGES is a multiple of four), LBO = one chunk; a K16 step = two 8-key groups = 2 KB further down.
    const uint32_t idesc_mn = idesc_full | (1u << 16);    // b_major = MN
    auto mma2 = [&](uint32_t T, bool first, bool last) {  // O += P(T) V(T)
      prof_add<PROF>(w_pf, mbar_wait(p_full, T & 1));
      if (first) prof_add<PROF>(w_oe, mbar_wait(o_empty, (it & 1) ^ 1));        // the epilogue of the previous item has drained O
      for (int half = 0; half < 2; ++half) {              // V.hi chunks, then V.lo chunks
        const int s0 = stage;
        prof_add<PROF>(w_rf, mbar_wait(r_full(s0), phase));
        prof_add<PROF>(w_rf, mbar_wait(r_full(s0 + 1), phase));
        tc_fence_after();
        if (leader) {
          if (p.v_chn) {
            // channel-major V (the memory read gets it that way): the classic K-major operand, 64 keys per chunk
            const uint64_t bd = make_kmajor_sw128_desc(sbase + OFF_RING + s0 * CHUNK_BYTES);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t b = bd + (uint64_t)((kk >> 2) * (CHUNK_BYTES >> 4) + 2 * (kk & 3));
              umma_bf16_ts(t_o, t_p + (uint32_t)(8 * kk), b, idesc_full, (first && half == 0 && kk == 0) ? 0u : 1u);
              if (half == 0) umma_bf16_ts(t_o, t_p + 64 + (uint32_t)(8 * kk), b, idesc_full, 1u);
            }
          } else {
            const uint64_t bd = make_mnmajor_sw128_desc(sbase + OFF_RING + s0 * CHUNK_BYTES, CHUNK_BYTES);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t b = bd + (uint64_t)(kk * (2048 >> 4));
              umma_bf16_ts(t_o, t_p + (uint32_t)(8 * kk), b, idesc_mn, (first && half == 0 && kk == 0) ? 0u : 1u);   // P.hi V.(hi|lo)
              if (half == 0) umma_bf16_ts(t_o, t_p + 64 + (uint32_t)(8 * kk), b, idesc_mn, 1u);                     // P.lo V.hi
            }
          }
          umma_commit(r_empty(s0));
          umma_commit(r_empty(s0 + 1));
          if (half == 1) {
            umma_commit(p_empty);
            if (last) umma_commit(o_full);
          }
        }
        __syncwarp();
        stage += 2;
        if (stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const AItem w = locate_item(ap, item, ns);
      const int kb = w.kb, ke = w.ke, nkt = w.nkt;
      idesc_tail = make_idesc(((w.nk - (nkt - 1) * TN) + 15) & ~15);
      prof_add<PROF>(w_qf, mbar_wait(q_full, it & 1));
      tc_fence_after();
      // Q is read by UMMA-1 only: it is handed back as soon as the LAST S tile of the item has been issued, so the
      // producer brings in the next item's Q tile and first K chunks under the last two UMMA-2 groups
      mma1(tile, kb == nkt - 1);
      if (kb + 1 == ke && leader) umma_commit(q_empty);
      __syncwarp();
      for (int t = kb; t < ke; ++t) {
        if (t + 1 < ke) {
          mma1(tile + (t - kb) + 1, t + 1 == nkt - 1);
          if (t + 2 == ke && leader) umma_commit(q_empty);
          __syncwarp();
        }
        mma2(tile + (t - kb), t == kb, t == ke - 1);
      }
      tile += ke - kb;
    }
    if (PROF && ap.prof && leader) {
      unsigned long long* o = ap.prof + blockIdx.x * 16;
      o[8] = w_se; o[9] = w_rf; o[10] = w_pf; o[11] = w_oe; o[12] = w_qf; o[13] = clock64() - t_begin;
    }
    __syncwarp();
  } else {
    // ===================== softmax + epilogue =====================
    const int quarter = warp & 3, part = warp >> 2;        // TMEM lane quarter; 32-column part of every key tile
    const uint32_t xm_u32 = sbase + OFF_XM, xl_u32 = sbase + OFF_XL;
    const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const float c2 = LOG2E * ap.inv_sqrt_c;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int cb = part * 32;                              // first column (within the tile) of my chunk
    uint32_t tile = 0, it = 0;
    long long w_sf = 0, t_ld = 0, t_x = 0, t_e = 0, w_pe = 0, t_st = 0, w_of = 0;
    const long long t_begin = PROF ? clock64() : 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const AItem w = locate_item(ap, item, ns);
      const int ks = w.ks, rt = w.rt, prob = w.prob, kb = w.kb, ke = w.ke, nq = w.nq, nk = w.nk;
      const int wprob = prob + ap.seg_prob0[w.seg];        // window map: block index counted over all sets
      const int row = rt * TM + quarter * 32 + lane;
      const bool row_ok = row < nq;
      float mref = -INFINITY, l = 0.f;                     // reference exponent (log2 units), my part of the row sum
      for (int kt = kb; kt < ke; ++kt, ++tile) {
        const int buf = tile & 1;
        prof_add<PROF>(w_sf, mbar_wait(s_full(buf), (tile >> 1) & 1));
        long long c0 = PROF ? clock64() : 0;
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32_async(tmem_base + lane_base + (uint32_t)(buf * TN + cb), r);
        tmem_wait(r);
        tc_fence_before();
        mbar_arrive(s_empty(buf));                          // my part of the S tile is in registers
        if (PROF) { const long long c1 = clock64(); t_ld += c1 - c0; c0 = c1; }
        const int nv = nk - (kt * TN + cb);                 // valid columns of my part (warp-uniform)
        if (nv < 32) {                                      // key tail: masked scores = -inf -> max ignores them, P = 0
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = (i < nv) ? r[i] : 0xff800000u;
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * c2;
        // row maximum of this key tile: the four parts of a row live in four warps of the same lane quarter
        const uint32_t slot = xm_u32 + (uint32_t)((tile & 1) * (NMATH * 32 * 4));
        sts32(slot + (uint32_t)((warp * 32 + lane) * 4), mx);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
        const uint32_t sq = slot + (uint32_t)((quarter * 32 + lane) * 4);
        const float m_tile = fmaxf(fmaxf(lds32(sq), lds32(sq + 4 * 128)), fmaxf(lds32(sq + 8 * 128), lds32(sq + 12 * 128)));
        if (PROF) { const long long c1 = clock64(); t_x += c1 - c0; c0 = c1; }
        float f = 1.f;
        const bool grow = m_tile > mref + TAU;              // first tile: mref = -inf
        if (grow) {
          f = ex2f(mref - m_tile);                          // first tile: 0
          mref = m_tile;
          l *= f;
        }
        uint32_t phi[16], plo[16];
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float e0 = ex2f(fmaf(__uint_as_float(r[2 * q]), c2, -mref));
          float e1 = ex2f(fmaf(__uint_as_float(r[2 * q + 1]), c2, -mref));
          ls[q & 3] += e0 + e1;
          split_bf16x2_alu(e0, e1, phi[q], plo[q]);
        }
        l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        // the P buffer is free, and O is quiescent, once UMMA-2 of the previous tile has retired
        if (PROF) { const long long c1 = clock64(); t_e += c1 - c0; }
        prof_add<PROF>(w_pe, mbar_wait(p_empty, (tile & 1) ^ 1));
        if (PROF) c0 = clock64();
        tc_fence_after();
        if (kt != kb && __any_sync(0xffffffffu, grow)) {    // rare: rescale my 32 columns of the O accumulator
          uint32_t o[32];
          tmem_ld32_async(tmem_base + lane_base + COL_O + (uint32_t)cb, o);
          tmem_wait(o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
          tmem_st32(tmem_base + lane_base + COL_O + (uint32_t)cb, o);
        }
        const uint32_t t_phi = tmem_base + lane_base + COL_P + (uint32_t)(part * 16);
        tmem_st16(t_phi, phi);
        tmem_st16(t_phi + 64, plo);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(p_full);
        if (PROF) t_st += clock64() - c0;
      }
      // ---- epilogue: row sum across the four parts, O / l -> global
      sts32(xl_u32 + (uint32_t)((warp * 32 + lane) * 4), l);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
      const uint32_t lq = xl_u32 + (uint32_t)((quarter * 32 + lane) * 4);
      const float l_row = (lds32(lq) + lds32(lq + 4 * 128)) + (lds32(lq + 8 * 128) + lds32(lq + 12 * 128));
      prof_add<PROF>(w_of, mbar_wait(o_full, it & 1));
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32_async(tmem_base + lane_base + COL_O + (uint32_t)cb, r);
        tmem_wait(r);
        tc_fence_before();
        mbar_arrive(o_empty);
        float* OUT;
        float sc;
        if (ns == 1) {
          OUT = p.out + (size_t)prob * p.out_stride_b;
          sc = 1.0f / l_row;
          if (part == 0 && row_ok && p.lse != nullptr)
            p.lse[p.win.enabled ? p.win.pixel(wprob, row) : (size_t)prob * nq + row] = mref * LN2 + logf(l_row);
        } else {
          OUT = p.part_o + ((size_t)ks * p.nb + prob) * nq * 128;
          sc = 1.0f;
          if (part == 0 && row_ok) p.part_ml[((size_t)ks * p.nb + prob) * nq + row] = make_float2(mref, l_row);
        }
        if (row_ok) {
          if (p.out_layout == EMIP_LAYOUT_NC) {
            size_t orow = (size_t)row;
            if (p.win.enabled) {                          // scatter: token `row` of block (prob / B) of image (prob % B)
              OUT = p.out;
              orow = p.win.pixel(wprob, row);
            }
            if (p.out != nullptr || ns != 1) {
            float4* dst = reinterpret_cast<float4*>(OUT + orow * 128 + cb);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              dst[q] = make_float4(__uint_as_float(r[4 * q]) * sc, __uint_as_float(r[4 * q + 1]) * sc,
                                   __uint_as_float(r[4 * q + 2]) * sc, __uint_as_float(r[4 * q + 3]) * sc);
            }
            if (p.out_hi != nullptr && ns == 1) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const float v0 = __uint_as_float(r[2 * q]) * sc, v1 = __uint_as_float(r[2 * q + 1]) * sc;
                const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                hi[q] = *reinterpret_cast<const uint32_t*>(&h);
                const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - __uint_as_float(hi[q] << 16), v1 - __uint_as_float(hi[q] & 0xffff0000u));
                lo[q] = *reinterpret_cast<const uint32_t*>(&l);
              }
              uint4* dh = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out_hi) + orow * p.out_ld + cb);
              uint4* dl = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out_lo) + orow * p.out_ld + cb);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                dh[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                dl[q] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) OUT[(size_t)(cb + i) * nq + row] = __uint_as_float(r[i]) * sc;
          }
        }
      }
    }
    if (PROF && ap.prof && threadIdx.x == 0) {
      unsigned long long* o = ap.prof + blockIdx.x * 16;
      o[0] = w_sf; o[1] = t_ld; o[2] = t_x; o[3] = t_e; o[4] = w_pe; o[5] = t_st; o[6] = w_of; o[7] = clock64() - t_begin;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NMATH + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// out = sum_s O_s 2^{m_s - M} / sum_s l_s 2^{m_s - M},  M = max_s m_s  (the key splits of one row)
__global__ void __launch_bounds__(256)
attn_merge_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml, float* __restrict__ out,
                  long long out_stride_b, float* __restrict__ lse, int ns, int nb, int nq, int layout) {
  const int b = blockIdx.y;
  const size_t per = (size_t)128 * nq;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
    const int row = layout == EMIP_LAYOUT_NC ? (int)(i >> 7) : (int)(i % nq);
    float M = -INFINITY;
    for (int s = 0; s < ns; ++s) M = fmaxf(M, __ldg(&part_ml[((size_t)s * nb + b) * nq + row]).x);
    float acc = 0.f, wsum = 0.f;
    for (int s = 0; s < ns; ++s) {
      const float2 ml = __ldg(&part_ml[((size_t)s * nb + b) * nq + row]);
      const float w = exp2f(ml.x - M);
      wsum = fmaf(ml.y, w, wsum);
      acc = fmaf(__ldcs(part_o + ((size_t)s * nb + b) * per + i), w, acc);
    }
    out[(size_t)b * out_stride_b + i] = acc / wsum;
    const bool first_of_row = layout == EMIP_LAYOUT_NC ? (i & 127) == 0 : i < (size_t)nq;
    if (lse != nullptr && first_of_row) lse[(size_t)b * nq + row] = M * 0.6931471805599453f + logf(wsum);
  }
}

}  // namespace

static unsigned long long* g_attn_prof = nullptr;
// Diagnostics: device buffer of gridDim.x * 16 counters filled by the next launches (NULL switches it off).
extern "C" void emip_attn_tc_set_profile_buffer(unsigned long long* dev_buf) { g_attn_prof = dev_buf; }
unsigned long long* attn_tc_profile_buffer() { return g_attn_prof; }

bool attn_tc_supported(int nq, int nk, int c) { return c == 128 && nq >= 1 && nk >= 1; }

int attn_tc_fwd(const AttnTcArgs& a, cudaStream_t st) {
  if (a.nb == 0 || a.nq == 0) return EMIP_OK;
  if (!attn_tc_supported(a.nq, a.nk, 128)) { emip_set_error("attn_tc_fwd: unsupported shape"); return EMIP_ENOSYS; }
  const int nkt = (a.nk + TN - 1) / TN;
  if (a.ksplit > nkt) { emip_set_error("attn_tc_fwd: ksplit %d exceeds the %d key tiles", a.ksplit, nkt); return EMIP_EINVAL; }
  if (a.win.enabled && (a.ksplit > 1 || a.out_layout != EMIP_LAYOUT_NC)) { emip_set_error("attn_tc_fwd: the window scatter needs ksplit == 1 and the NC layout"); return EMIP_EINVAL; }
  if (a.ksplit > 1 && (a.part_o == nullptr || a.part_ml == nullptr)) { emip_set_error("attn_tc_fwd: ksplit needs partial buffers"); return EMIP_EINVAL; }
  if (a.n_more < 0 || a.n_more > MAXSEG - 1 || (a.n_more > 0 && (!a.win.enabled || a.ksplit > 1 || a.v_chn))) {
    emip_set_error("attn_tc_fwd: further problem sets need the window scatter, ksplit == 1 and token-major values");
    return EMIP_EINVAL;
  }
  AMaps maps;
  AParams ap;
  ap.a = a;
  ap.inv_sqrt_c = 1.0f / a.sqrt_c;
  ap.prof = g_attn_prof;
  ap.nseg = 1 + a.n_more;
  int rc;
  long long grid = 0;
  for (int s = 0; s < MAXSEG; ++s) {
    const bool used = s < ap.nseg;
    const void* q = s == 0 ? a.q_split : used ? a.more[s - 1].q_split : a.q_split;
    const void* k = s == 0 ? a.k_split : used ? a.more[s - 1].k_split : a.k_split;
    const void* v = s == 0 ? a.v_split : used ? a.more[s - 1].v_split : a.v_split;
    const int nb = s == 0 ? a.nb : used ? a.more[s - 1].nb : a.nb, nq = s == 0 ? a.nq : used ? a.more[s - 1].nq : a.nq;
    const int nk = s == 0 ? a.nk : used ? a.more[s - 1].nk : a.nk;
    if (used && (nb < 1 || !attn_tc_supported(nq, nk, 128))) { emip_set_error("attn_tc_fwd: unsupported shape in problem set %d", s); return EMIP_ENOSYS; }
    if ((rc = make_bf16_map(&maps.q[s], q, 256, (uint64_t)nq, (uint64_t)nb, 512, (uint64_t)nq * 512))) return rc;
    if ((rc = make_bf16_map(&maps.k[s], k, 256, (uint64_t)nk, (uint64_t)nb, 512, (uint64_t)nk * 512))) return rc;
    if (a.v_chn) {
      const uint64_t ld = ((uint64_t)nk + 7) / 8 * 8;     // = pair_bwd_tc_chn_ld
      if ((rc = make_bf16_map(&maps.v[s], v, (uint64_t)nk, 256, (uint64_t)nb, ld * 2, ld * 2 * 256))) return rc;
    } else if ((rc = make_bf16_map(&maps.v[s], v, 256, (uint64_t)nk, (uint64_t)nb, 512, (uint64_t)nk * 512))) return rc;
    const long long items = used ? (long long)nb * ((nq + TM - 1) / TM) * (a.ksplit > 1 ? a.ksplit : 1) : 0;
    if (items > 0x3fffffffLL) { emip_set_error("attn_tc_fwd: too many items"); return EMIP_EINVAL; }
    ap.seg_items[s] = (int)items;
    ap.seg_nq[s] = nq;
    ap.seg_nk[s] = nk;
    ap.seg_prob0[s] = (s > 0 && used) ? a.more[s - 1].blk0 * a.win.B : 0;
    grid += items;
  }
  if (int rc__ = emip_func_max_smem((const void*)(attn_fwd_tc_kernel<false>), SMEM_BYTES)) return rc__;
  if (int rc__ = emip_func_max_smem((const void*)(attn_fwd_tc_kernel<true>), SMEM_BYTES)) return rc__;
  if (grid > emip_num_sms()) grid = emip_num_sms();
  if (ap.prof) attn_fwd_tc_kernel<true><<<(int)grid, NTHREADS, SMEM_BYTES, st>>>(maps, ap);   // diagnostics build
  else EMIP_CUDA(emip_launch_pdl(attn_fwd_tc_kernel<false>, dim3((unsigned)grid), dim3(NTHREADS), (size_t)SMEM_BYTES, st, maps, ap));
  EMIP_CHECK_LAUNCH("attn_tc_fwd");
  return EMIP_OK;
}

int attn_tc_merge(const AttnTcArgs& a, cudaStream_t st) {
  if (a.nb == 0 || a.nq == 0 || a.ksplit <= 1) return EMIP_OK;
  attn_merge_kernel<<<dim3(128, a.nb), 256, 0, st>>>(a.part_o, a.part_ml, a.out, a.out_stride_b, a.lse, a.ksplit, a.nb, a.nq,
                                                    a.out_layout);
  EMIP_CHECK_LAUNCH("attn_tc_merge");
  return EMIP_OK;
}
