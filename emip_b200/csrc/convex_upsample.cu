// f4: convex x8 upsampling of the coarse flow -- forward and backward.
//
// Replaces reference model/EMIP_short/motion/gmflow/gmflow.py:64-77, the part of GMFlow.upsample_flow after the
// upsampler convolution: view [B,1,9,K,K,h,w] -> softmax over the 9 taps -> F.unfold(K*flow, 3x3, padding 1) ->
// weighted sum -> permute/reshape to [B,2,K*h,K*w].  The reference makes five passes over the 4.5 MB/sample mask
// (view, softmax read+write, product, sum) and materialises the [B,2,9,h,w] unfold; here the mask is read exactly
// once, the 3 coarse flow rows live in shared memory and the output rows are written contiguously.
// HBM-bound: algorithmic bytes per sample = (576 + 2) h w 4 in + 2 * 64 h w 4 out (5.5 MB at 44 x 44).
//
// One CTA = one coarse row y of one sample.  Threads own a fixed coarse column x (lanes = consecutive x, so every mask
// row segment is read coalesced) and loop over the 64 sub-positions; the fine output tile [2][8][8w] is staged in
// shared memory and stored as 16 contiguous rows.  Backward: the same loop recomputes the softmax, writes dmask
// directly, and reduces the flow gradient deterministically (per-thread partial sums over the sub-positions, fixed-
// order combine in shared memory, one [9][2][w] partial per coarse row; a second kernel gathers the 9 partials).
#include "common.cuh"
#include "../../include/emip_b200.h"

namespace {

constexpr int K = 8;            // upsample factor (configs.yaml: upsample_factor 8)
constexpr int KK = K * K;
constexpr int NT = 256;

__device__ __forceinline__ void load_flow_rows(float* sf, const float* __restrict__ flow, int b, int y, int h, int w) {
  // sf[c][r][x+1] = K * flow[b][c][y-1+r][x], zero padded: [2][3][w+2]
  const int wp = w + 2;
  for (int i = threadIdx.x; i < 2 * 3 * wp; i += NT) {
    const int c = i / (3 * wp), r = (i / wp) % 3, xx = i % wp - 1, yy = y - 1 + r;
    float v = 0.f;
    if (xx >= 0 && xx < w && yy >= 0 && yy < h) v = (float)K * __ldg(flow + (((size_t)b * 2 + c) * h + yy) * w + xx);
    sf[i] = v;
  }
}

__device__ __forceinline__ void softmax9(const float* __restrict__ mp, size_t tap_stride, float (&p)[9]) {
  float m = -INFINITY;
#pragma unroll
  for (int n = 0; n < 9; ++n) { p[n] = __ldcs(mp + n * tap_stride); m = fmaxf(m, p[n]); }   // streamed: read once
  float l = 0.f;
#pragma unroll
  for (int n = 0; n < 9; ++n) { p[n] = __expf(p[n] - m); l += p[n]; }
  const float inv = 1.0f / l;
#pragma unroll
  for (int n = 0; n < 9; ++n) p[n] *= inv;
}

__global__ void __launch_bounds__(NT)
convex_up_fwd_kernel(const float* __restrict__ flow, const float* __restrict__ mask, float* __restrict__ out, int h, int w) {
  extern __shared__ float sm[];
  const int y = blockIdx.x, b = blockIdx.y, wp = w + 2, W8 = K * w;
  float* sf = sm;                          // [2][3][w+2]
  float* so = sm + 2 * 3 * wp;             // [2][8][8w]
  load_flow_rows(sf, flow, b, y, h, w);
  __syncthreads();
  const size_t plane = (size_t)h * w, tap_stride = (size_t)KK * plane;
  const float* mb = mask + (size_t)b * 9 * KK * plane + (size_t)y * w;
  for (int idx = threadIdx.x; idx < KK * w; idx += NT) {
    const int s = idx / w, x = idx - s * w;
    float p[9];
    softmax9(mb + (size_t)s * plane + x, tap_stride, p);
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int n = 0; n < 9; ++n) {
      const int r = n / 3, dx = n % 3;
      o0 = fmaf(p[n], sf[(0 * 3 + r) * wp + x + dx], o0);
      o1 = fmaf(p[n], sf[(1 * 3 + r) * wp + x + dx], o1);
    }
    const int ky = s / K, kx = s % K;
    so[(0 * K + ky) * W8 + K * x + kx] = o0;
    so[(1 * K + ky) * W8 + K * x + kx] = o1;
  }
  __syncthreads();
  // 16 contiguous output rows of 8w floats
  for (int i = threadIdx.x; i < 2 * K * W8; i += NT) {
    const int c = i / (K * W8), ky = (i / W8) % K, xx = i % W8;
    __stcs(out + (((size_t)b * 2 + c) * (K * h) + (size_t)K * y + ky) * W8 + xx, so[i]);
  }
}

// dmask written directly; flow-gradient partials part[b][y][n][c][x] = sum_s p_n(s) dout[c][s]  (x K applied later)
__global__ void __launch_bounds__(NT)
convex_up_bwd_kernel(const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ dout,
                     float* __restrict__ dmask, float* __restrict__ part, int h, int w, int groups) {
  extern __shared__ float sm[];
  const int y = blockIdx.x, b = blockIdx.y, wp = w + 2, W8 = K * w;
  float* sf = sm;                          // [2][3][w+2]
  float* sg = sf + 2 * 3 * wp;             // dout tile [2][8][8w]
  float* sp = sg + 2 * K * W8;             // [groups][18][w]
  load_flow_rows(sf, flow, b, y, h, w);
  for (int i = threadIdx.x; i < 2 * K * W8; i += NT) {
    const int c = i / (K * W8), ky = (i / W8) % K, xx = i % W8;
    sg[i] = __ldcs(dout + (((size_t)b * 2 + c) * (K * h) + (size_t)K * y + ky) * W8 + xx);
  }
  __syncthreads();
  const size_t plane = (size_t)h * w, tap_stride = (size_t)KK * plane;
  const float* mb = mask + (size_t)b * 9 * KK * plane + (size_t)y * w;
  float* db = dmask + (size_t)b * 9 * KK * plane + (size_t)y * w;
  const int g = threadIdx.x / w, x = threadIdx.x - g * w;     // threads beyond groups * w idle in the main loop
  float acc[18];
#pragma unroll
  for (int i = 0; i < 18; ++i) acc[i] = 0.f;
  if (g < groups) {
    for (int s = g; s < KK; s += groups) {
      float p[9];
      softmax9(mb + (size_t)s * plane + x, tap_stride, p);
      const int ky = s / K, kx = s % K;
      const float g0 = sg[(0 * K + ky) * W8 + K * x + kx], g1 = sg[(1 * K + ky) * W8 + K * x + kx];
      float dp[9], dot = 0.f;
#pragma unroll
      for (int n = 0; n < 9; ++n) {
        const int r = n / 3, dx = n % 3;
        dp[n] = g0 * sf[(0 * 3 + r) * wp + x + dx] + g1 * sf[(1 * 3 + r) * wp + x + dx];
        dot = fmaf(p[n], dp[n], dot);
        acc[2 * n] = fmaf(p[n], g0, acc[2 * n]);
        acc[2 * n + 1] = fmaf(p[n], g1, acc[2 * n + 1]);
      }
#pragma unroll
      for (int n = 0; n < 9; ++n) db[(size_t)s * plane + n * tap_stride + x] = p[n] * (dp[n] - dot);   // softmax backward
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) sp[(g * 18 + i) * w + x] = acc[i];
  }
  __syncthreads();
  float* pb = part + ((size_t)b * h + y) * 18 * w;
  for (int i = threadIdx.x; i < 18 * w; i += NT) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += sp[gg * 18 * w + i];          // fixed order: deterministic
    pb[i] = s;
  }
}

// dflow[b][c][y'][x'] = K * sum_n part[b][y'-dy_n][n][c][x'-dx_n]   (tap n of coarse pixel (y,x) reads (y+dy_n, x+dx_n))
__global__ void convex_up_dflow_kernel(const float* __restrict__ part, float* __restrict__ dflow, int B, int h, int w) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * 2 * h * w) return;
  const int xq = (int)(i % w), yq = (int)((i / w) % h), c = (int)((i / ((long long)w * h)) % 2), b = (int)(i / ((long long)2 * w * h));
  float s = 0.f;
#pragma unroll
  for (int n = 0; n < 9; ++n) {
    const int y = yq - (n / 3 - 1), x = xq - (n % 3 - 1);
    if (y >= 0 && y < h && x >= 0 && x < w) s += __ldg(part + (((size_t)b * h + y) * 18 + 2 * n + c) * w + x);
  }
  dflow[i] = (float)K * s;
}

size_t fwd_smem(int w) { return sizeof(float) * (2 * 3 * (w + 2) + 2 * K * K * w); }
int bwd_groups(int w) { int g = NT / w; return g < 1 ? 1 : (g > KK ? KK : g); }
size_t bwd_smem(int w) { return sizeof(float) * (2 * 3 * (w + 2) + 2 * K * K * w + (size_t)bwd_groups(w) * 18 * w); }

int check(const char* who, int B, int h, int w, int k) {
  EMIP_CHECK_ARG(B >= 0 && h > 0 && w > 0, "%s: bad shape B=%d h=%d w=%d", who, B, h, w);
  if (k != K) { emip_set_error("%s: upsample factor %d unsupported (kernels are built for the model's 8)", who, k); return EMIP_ENOSYS; }
  if (w > NT || bwd_smem(w) > 200 * 1024) { emip_set_error("%s: coarse width %d too large for the row-resident kernels", who, w); return EMIP_ENOSYS; }
  return EMIP_OK;
}

}  // namespace

extern "C" size_t emip_convex_upsample_workspace(int B, int h, int w) {
  if (B < 0 || h <= 0 || w <= 0) return 0;
  return emip_align_up(sizeof(float) * (size_t)B * h * 18 * w, 256);
}

extern "C" int emip_convex_upsample_fwd(const float* flow, const float* mask, float* out, int B, int h, int w, int k,
                                        void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(flow && mask && out, "convex_upsample_fwd: null pointer");
  int rc = check("convex_upsample_fwd", B, h, w, k);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc__ = emip_func_max_smem((const void*)(convex_up_fwd_kernel), 200 * 1024)) return rc__;
  if (int rc__ = emip_func_max_smem((const void*)(convex_up_bwd_kernel), 200 * 1024)) return rc__;
  convex_up_fwd_kernel<<<dim3(h, B), NT, fwd_smem(w), st>>>(flow, mask, out, h, w);
  EMIP_CHECK_LAUNCH("convex_up_fwd");
  return EMIP_OK;
}

extern "C" int emip_convex_upsample_bwd(const float* flow, const float* mask, const float* dout, float* dflow, float* dmask,
                                        void* workspace, size_t ws_bytes, int B, int h, int w, int k, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(flow && mask && dout && dflow && dmask && workspace, "convex_upsample_bwd: null pointer");
  int rc = check("convex_upsample_bwd", B, h, w, k);
  if (rc) return rc;
  if (ws_bytes < emip_convex_upsample_workspace(B, h, w)) { emip_set_error("convex_upsample_bwd: workspace too small"); return EMIP_ENOMEM; }
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc__ = emip_func_max_smem((const void*)(convex_up_fwd_kernel), 200 * 1024)) return rc__;
  if (int rc__ = emip_func_max_smem((const void*)(convex_up_bwd_kernel), 200 * 1024)) return rc__;
  float* part = static_cast<float*>(workspace);
  convex_up_bwd_kernel<<<dim3(h, B), NT, bwd_smem(w), st>>>(flow, mask, dout, dmask, part, h, w, bwd_groups(w));
  EMIP_CHECK_LAUNCH("convex_up_bwd");
  const long long n = (long long)B * 2 * h * w;
  convex_up_dflow_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, dflow, B, h, w);
  EMIP_CHECK_LAUNCH("convex_up_dflow");
  return EMIP_OK;
}
