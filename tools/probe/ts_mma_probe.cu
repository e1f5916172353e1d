// Probe: TS-mode UMMA (A operand in TMEM, written by tcgen05.cp) vs SS-mode on the same data.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ float aval(int r, int k) { return (float)((r * 3 + k * 5) % 7 - 3); }
__device__ float bval(int n, int k) { return (float)((n * 2 + k * 3) % 5 - 2); }

__global__ void probe(float* out, int a_step) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 16384;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * 64; i += blockDim.x) {
    int r = i / 64, k = i % 64, j = k / 8, e = k % 8;
    *(__nv_bfloat16*)(sA + r * 128 + ((j ^ (r & 7)) * 16) + e * 2) = __float2bfloat16(aval(r, k));
    *(__nv_bfloat16*)(sB + r * 128 + ((j ^ (r & 7)) * 16) + e * 2) = __float2bfloat16(bval(r, k));
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_slot;
  if (tid == 0) {
    const uint64_t ad = make_desc(smem_u32(sA)), bd = make_desc(smem_u32(sB));
    const uint32_t idesc = make_idesc(128);
    // SS reference into columns [0,128)
    for (int k = 0; k < 4; ++k) {
      uint32_t acc = k ? 1u : 0u;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tb), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc) : "memory");
    }
    // A -> TMEM columns [256, 288)
    for (int k = 0; k < 4; ++k)
      asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tb + 256 + 8 * k), "l"(ad + 2 * k) : "memory");
    // TS into columns [128,256)
    for (int k = 0; k < 4; ++k) {
      uint32_t acc = k ? 1u : 0u;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                   ::"r"(tb + 128), "r"(tb + 256 + a_step * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tb + ((uint32_t)((warp & 3) * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 256 + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

static float ha(int r, int k) { return (float)((r * 3 + k * 5) % 7 - 3); }
static float hb(int n, int k) { return (float)((n * 2 + k * 3) % 5 - 2); }
int main() {
  float* d;
  cudaMalloc(&d, 128 * 256 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  static float h[128 * 256];
  for (int a_step : {8, 16, 4}) {
    cudaMemset(d, 0, 128 * 256 * 4);
    probe<<<1, 128, 40 * 1024>>>(d, a_step);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad_ss = 0, bad_ts = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < 128; ++n) {
        float ref = 0;
        for (int k = 0; k < 64; ++k) ref += ha(r, k) * hb(n, k);
        if (h[r * 256 + n] != ref) ++bad_ss;
        if (h[r * 256 + 128 + n] != ref) ++bad_ts;
      }
    printf("a_step %2d: %s  SS mismatches %d  TS mismatches %d   D_ss[5][0..3] = %g %g %g %g  D_ts[5][0..3] = %g %g %g %g\n", a_step,
           cudaGetErrorString(e), bad_ss, bad_ts, h[5 * 256], h[5 * 256 + 1], h[5 * 256 + 2], h[5 * 256 + 3], h[5 * 256 + 128],
           h[5 * 256 + 129], h[5 * 256 + 130], h[5 * 256 + 131]);
  }
  return 0;
}
