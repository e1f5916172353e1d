// f3: backward-flow occlusion mask of the photometric loss.
//
// Replaces reference loss/warp_utils.py:106-112 get_occu_mask_backward(flow21, th) = (clamp(splat, 0, 1) < th), where
// splat = get_corresponding_map(base_grid + flow21) (warp_utils.py:26-80): every pixel deposits its four bilinear
// weights at the integer neighbours of its target position; taps outside the image are dropped.  Not differentiable
// (a comparison), so forward only.  The reference builds four [B,N] index / value / invalid tensors, concatenates
// them and calls scatter_add_ (itself atomic on the GPU); here one pass reads the flow through its channel-slice
// strides (loss_flow.py:95-96) and issues the four red.global.add directly, a second pass thresholds.
// HBM-bound: 2 planes in, 1 plane of L2-resident atomics, 1 plane out.
#include "common.cuh"
#include "../../include/emip_b200.h"

namespace {

__global__ void __launch_bounds__(256)
splat_kernel(const float* __restrict__ flow, float* __restrict__ acc, int H, int W, long long fsb, long long fsc) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const float* fp = flow + (long long)b * fsb + (long long)y * W + x;
  const float px = (float)x + __ldcs(fp), py = (float)y + __ldcs(fp + fsc);
  const float x1 = floorf(px), y1 = floorf(py);
  float* ab = acc + (size_t)b * H * W;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
#pragma unroll
  for (int t = 0; t < 4; ++t) {                       // (ceil,ceil) (ceil,floor) (floor,ceil) (floor,floor)
    const float xc = x1 + ((t & 2) ? 0.f : 1.f), yc = y1 + ((t & 1) ? 0.f : 1.f);
    if (xc >= 0.f && xc <= wm1 && yc >= 0.f && yc <= hm1) {
      const float v = (1.0f - fabsf(px - xc)) * (1.0f - fabsf(py - yc));      // warp_utils.py:67-71
      atomicAdd(ab + (int)yc * W + (int)xc, v);
    }
  }
}

__global__ void threshold_kernel(const float* __restrict__ acc, float* __restrict__ mask, long long n, float th) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = fminf(fmaxf(acc[i], 0.f), 1.f);     // warp_utils.py:111
  mask[i] = v < th ? 1.0f : 0.0f;
}

}  // namespace

extern "C" size_t emip_occu_mask_workspace(int B, int H, int W) {
  if (B < 0 || H <= 0 || W <= 0) return 0;
  return emip_align_up(sizeof(float) * (size_t)B * H * W, 256);
}

extern "C" int emip_occu_mask_backward(const float* flow21, float* mask, void* workspace, size_t ws_bytes, int B, int H, int W,
                                       long long flow_stride_b, long long flow_stride_c, float th, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(flow21 && mask && workspace, "occu_mask_backward: null pointer");
  EMIP_CHECK_ARG(B >= 0 && H > 0 && W > 0 && B <= 65535, "occu_mask_backward: bad shape B=%d H=%d W=%d", B, H, W);
  if (ws_bytes < emip_occu_mask_workspace(B, H, W)) { emip_set_error("occu_mask_backward: workspace too small"); return EMIP_ENOMEM; }
  cudaStream_t st = (cudaStream_t)stream;
  float* acc = static_cast<float*>(workspace);
  const long long n = (long long)B * H * W;
  EMIP_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * n, st));
  splat_kernel<<<dim3((W + 31) / 32, (H + 7) / 8, B), 256, 0, st>>>(flow21, acc, H, W, flow_stride_b, flow_stride_c);
  EMIP_CHECK_LAUNCH("occu splat");
  threshold_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(acc, mask, n, th);
  EMIP_CHECK_LAUNCH("occu threshold");
  return EMIP_OK;
}
