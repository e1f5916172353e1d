// Probe: where do the bytes of a K-major SW128 smem tile land in TMEM after tcgen05.cp.128x256b?
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_cp_probe tmem_cp_probe.cu && ./tmem_cp_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

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

__global__ void probe(uint32_t* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // logical element: row r (0..127), 16-byte chunk j (0..7), word w (0..3)  ->  value r*1000 + j*10 + w
  for (int i = tid; i < 128 * 32; i += blockDim.x) {
    int r = i / 32, q = i % 32, j = q / 4, w = q % 4;
    uint32_t* dst = (uint32_t*)(smem + r * 128 + ((j ^ (r & 7)) * 16) + w * 4);
    *dst = r * 1000 + j * 10 + w;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_slot;
  if (tid == 0) {
    uint64_t sd = make_desc(smem_u32(smem));
    for (int k = 0; k < 4; ++k)
      asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tb + 8 * k), "l"(sd + 2 * k) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  const uint32_t taddr = tb + ((uint32_t)((warp & 3) * 32) << 16);
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
  for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 32 + i] = r[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tb) : "memory");
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 128 * 32 * 4);
  cudaMemset(d, 0xff, 128 * 32 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
  probe<<<1, 128, 20 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static uint32_t h[128 * 32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int lane : {0, 1, 2, 7, 8, 9, 31, 32, 33, 64, 127}) {
    printf("lane %3d:", lane);
    for (int c = 0; c < 32; ++c) printf(" %d", h[lane * 32 + c]);
    printf("\n");
  }
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 32; ++c) {
      int k = c / 8, q = c % 8;
      uint32_t want = r * 1000 + (2 * k + q / 4) * 10 + q % 4;
      if (h[r * 32 + c] != want) ++bad;
    }
  printf("mismatches vs (lane=row, col=8k+q <- chunk 2k+q/4 word q%%4): %d of 4096\n", bad);
  return 0;
}
