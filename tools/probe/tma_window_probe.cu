// Probe: how fast does the TMA unit deliver small fp32 windows (box W x H x 3 planes of a [B*3, 352, 352] image tensor)
// into shared memory?  One thread per CTA keeps DEPTH window requests in flight and times the whole stream.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_window_probe tma_window_probe.cu -lcuda && ./tma_window_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int DEPTH_MAX = 8;

__global__ void __launch_bounds__(128)
probe(const __grid_constant__ CUtensorMap map, int box_bytes, int depth, int n_windows, int H, int W, int bw, int bh, int planes3,
      int xalign, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 127) & ~(uintptr_t)127);
  __shared__ __align__(8) unsigned long long bars[DEPTH_MAX];
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH_MAX; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int stride = (box_bytes + 127) / 128 * 128;
  unsigned rng = blockIdx.x * 2654435761u + 12345u;
  auto issue = [&](int i) {
    rng = rng * 1664525u + 1013904223u;
    int x = (int)((rng >> 8) % (unsigned)(W - bw));
    int y = (int)((rng >> 4) % (unsigned)(H - bh));
    int p = (int)((rng >> 16) % (unsigned)planes3) * 3;
    x = x / xalign * xalign;
    const int s = i % depth;
    const uint32_t bar = smem_u32(&bars[s]), dst = smem_u32(smem + s * stride);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(&map), "r"(bar), "r"(x), "r"(y), "r"(p) : "memory");
  };
  const long long t0 = clock64();
  for (int i = 0; i < depth && i < n_windows; ++i) issue(i);
  for (int i = 0; i < n_windows; ++i) {
    const int s = i % depth;
    const uint32_t parity = (uint32_t)(i / depth) & 1u;
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(smem_u32(&bars[s])), "r"(parity) : "memory");
    if (i + depth < n_windows) issue(i + depth);
  }
  cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 64, H = 352, W = 352;
  printf("image tensor %d x 3 x %d x %d fp32 = %.0f MB\n", B, H, W, (double)B * 3 * H * W * 4 / 1e6);
  float* x;
  cudaMalloc(&x, (size_t)B * 3 * H * W * 4);
  cudaMemset(x, 0, (size_t)B * 3 * H * W * 4);
  long long* cyc;
  cudaMalloc(&cyc, 1024 * 8);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fn;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int shapes[][2] = {{40, 40}, {64, 40}, {80, 40}, {40, 72}, {64, 64}};
  for (auto& sh : shapes) {
    const int bw = sh[0], bh = sh[1];
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * 3};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 3};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int box_bytes = bw * bh * 12;
    for (int xalign : {4})
      for (int ctas : {1, 2})
        for (int depth : {1, 2, 4, 8}) {
          const int grid = 148 * ctas, n_windows = 200;
          const int smem = depth * ((box_bytes + 127) / 128 * 128) + 256;
          if (smem * ctas > 220 * 1024) continue;
          cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          probe<<<grid, 128, smem>>>(map, box_bytes, depth, n_windows, H, W, bw, bh, B, xalign, cyc);
          cudaEventRecord(e0);
          probe<<<grid, 128, smem>>>(map, box_bytes, depth, n_windows, H, W, bw, bh, B, xalign, cyc);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          const double total = (double)grid * n_windows * box_bytes;
          printf("box %3dx%2dx3 (%5d B) xalign %2d  ctas/SM %d depth %d : %7.1f us  %6.0f GB/s  %6.2f us/window/CTA\n", bw, bh,
                 box_bytes, xalign, ctas, depth, ms * 1e3, total / (ms * 1e-3) / 1e9, ms * 1e3 / n_windows);
        }
  }
  return 0;
}
