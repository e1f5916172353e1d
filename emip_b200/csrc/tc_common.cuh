// PTX wrappers shared by the tcgen05 kernels (match_tc.cu, pair_bwd_tc.cu): mbarrier, TMA, UMMA, TMEM.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace tc {

constexpr int TM = 128;                 // rows per tile (UMMA M)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU.
// Returns the cycles spent waiting (0 on the fast path); used by the optional role profile.
__device__ __forceinline__ long long mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();                     // try_wait itself suspends the warp for a while
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("emip match_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar,
             parity);
      __trap();
    }
  }
  return clock64() - t0;
}
// role-profile accumulation that compiles away in the production instantiation of a kernel
template <bool PROF>
__device__ __forceinline__ void prof_add(long long& acc, long long v) {
  if constexpr (PROF) acc += v;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset
  d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}
// MN-major, 128-byte swizzle (cute::UMMA make_umma_desc<Major::MN>, LayoutType::B128): rows of 128 B hold 64 consecutive
// M/N elements of one K index, 8 K-rows form a 1 KB swizzle atom (stride byte offset between atoms along K), the next
// 64 M/N elements start `lbo_bytes` further (leading byte offset).  This is what a {64 elements x rows} SWIZZLE_128B
// TMA box of a row-major [K][MN] tensor leaves in shared memory.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

#define R32_OUT(r)                                                                                                     \
  "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),          \
      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define R32_INOUT(r)                                                                                                   \
  "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),          \
      "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),           \
      "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),          \
      "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])

// Asynchronous TMEM load of 32 consecutive fp32 columns of this thread's lane; pair with tmem_wait().
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : R32_OUT(r)
      : "r"(taddr)
      : "memory");
}
// The registers are in/out operands so that no use of r[] can be scheduled above the wait.
__device__ __forceinline__ void tmem_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : R32_INOUT(r)::"memory");
}
// 16-column variants (half the registers of the x32 forms)
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Warp-private epilogue staging tile: 32 rows x 128 B, the 16-byte piece c (0..7) of row r at piece c ^ stg_swz(r).  Eight
// consecutive rows have eight different keys, so a row-per-lane 16-byte store is conflict-free (each quarter warp covers all 32
// banks); rows r, r + 1 (r even) differ in bit 2, so a quarter warp reading pieces 0..3 of both rows is conflict-free as well,
// and so is one reading all eight pieces of one row.
__device__ __forceinline__ uint32_t stg_swz(uint32_t r) { return ((r & 1u) << 2) | ((r >> 1) & 3u); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// exact-GELU x Phi(x) for the epilogue (where the instruction count is the critical path): Phi through the rational
// erfc form of Abramowitz & Stegun 7.1.26, erfc(z) = poly5(t) e^{-z^2}, t = 1 / (1 + 0.3275911 z), |error| <= 1.5e-7 --
// below the 2^-17 relative resolution of the bf16 hi | lo pair the value is stored as.  One MUFU.RCP + one MUFU.EX2 and
// ~12 FP32 instructions instead of erff's two-branch polynomial; no 1 + erf cancellation for negative x.
__device__ __forceinline__ float gelu_epi(float x) {
  // 0.5 folded into the polynomial; rcp / ex2 as single MUFU instructions (.ftz: t is in (0, 1], e^{-z^2} may flush to 0)
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.44269504088896340736f * z));
  float p = fmaf(0.5307027145f, t, -0.7265760135f);
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  const float g = p * t * e;                                   // 0.5 erfc(|x| / sqrt 2) = Phi(-|x|)
  return x * (x >= 0.f ? 1.f - g : g);
}

// Two elements at a time on the packed fp32 pipe (FFMA2 / FMUL2: sm_100 executes two fp32 FMAs per instruction): the GELU / split
// epilogue of the 256 -> 1024 layer is instruction-issue bound (r3i: 25 instructions per element on eight warps).
// gelu(x) = max(x, 0) - |x| Phi(-|x|): no select, no 1 - g.
typedef unsigned long long f32x2;   // two packed fp32 (FFMA2 / FMUL2 operands)
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void gelu_epi2(float& x0, float& x1) {
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  const f32x2 ax = pk2(a0, a1);
  const f32x2 z = mul2(ax, pk2(0.70710678118654752440f, 0.70710678118654752440f));
  float d0, d1, q0, q1, t0, t1, e0, e1;
  upk2(fma2(z, pk2(0.3275911f, 0.3275911f), pk2(1.f, 1.f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  upk2(mul2(mul2(z, pk2(-1.44269504088896340736f, -1.44269504088896340736f)), z), q0, q1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
  const f32x2 t = pk2(t0, t1);
  f32x2 p = fma2(pk2(0.5307027145f, 0.5307027145f), t, pk2(-0.7265760135f, -0.7265760135f));
  p = fma2(p, t, pk2(0.7107068705f, 0.7107068705f));
  p = fma2(p, t, pk2(-0.142248368f, -0.142248368f));
  p = fma2(p, t, pk2(0.127414796f, 0.127414796f));
  const f32x2 g = mul2(mul2(p, t), pk2(e0, e1));                         // Phi(-|x|) of both elements
  upk2(fma2(pk2(-a0, -a1), g, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), x0, x1);
}


// fp32 pair -> packed bf16 hi (round to nearest, ties away from zero) + packed bf16 lo (truncated residual) on the
// integer / FMA pipes only.  cvt.rn.bf16x2.f32 (F2FP) shares the 16-lanes-per-clock XU pipe with MUFU.EX2: in the
// softmax loops, which need that pipe for the exponentials, two conversions per pair doubled the time per element
// (attn_tc role profile r2b: 17 cycles per warp-element with F2FP, MUFU alone is 8).  |a - (hi + lo)| <= 2^-16 |a|.
__device__ __forceinline__ void split_bf16x2_alu(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ha = (__float_as_uint(a) + 0x8000u) & 0xffff0000u, hb = (__float_as_uint(b) + 0x8000u) & 0xffff0000u;
  const float la = a - __uint_as_float(ha), lb = b - __uint_as_float(hb);
  hi = __byte_perm(ha, hb, 0x7632);                                  // (a.hi16, b.hi16): first element in the low half
  lo = __byte_perm(__float_as_uint(la), __float_as_uint(lb), 0x7632);
}

// 16 packed 32-bit columns of this thread's TMEM lane (tcgen05.st, 32x32b shape)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// TS mode: A operand from tensor memory ([128 lanes x 8 columns] per K16 step, two bf16 per column).
// Checked against the SS form on identical data in tools/probe/ts_mma_probe.cu.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// 3-D bf16 tensor map, 128-byte swizzle, box = {64 elements, 128 rows, 1}, out-of-bounds reads as zero
inline int make_bf16_map(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                         uint64_t stride2_bytes) {
  EncodeTiledFn enc = get_encoder();
  if (enc == nullptr) {
    emip_set_error("cuTensorMapEncodeTiled not available from the driver");
    return EMIP_ENOSYS;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {64, 128, 1};
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

}  // namespace tc
